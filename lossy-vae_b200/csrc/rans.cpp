// Host entropy coder: quantised-CDF builder and rANS64 encoder/decoder.
//
// Restates the published behaviour of CompressAI (un-vendored dependency of the reference; call
// sites lvae/models/qarv/model.py:106-113,123-124): cpp_exts/ops/ops.cpp pmf_to_quantized_cdf,
// cpp_exts/rans/rans_interface.cpp {encode,decode}_with_indexes and ryg_rans' rans64.h -- 64-bit
// state, 32-bit renormalisation words, 16-bit probabilities, 4-bit bypass for out-of-table symbols.
// Byte-compatibility with a real CompressAI build is "parity unpinned" (absent offline); the
// testable properties are equality with the Python restatement in oracle/ and lossless round trips.
// The coder is serial per stream and thread-safe: callers run one stream per (image, layer).  Speed: the encoder uses
// ryg_rans' reciprocal-multiply symbol form (bit-identical to the division form), the decoder a 256-bucket
// cumulative -> symbol table per row, and the multi-stream entry points advance two streams per worker in lockstep.
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "lvae_b200.h"

namespace {
constexpr int kPrecision = 16;
constexpr int kBypassPrecision = 4;
constexpr int kMaxBypassVal = (1 << kBypassPrecision) - 1;
constexpr uint64_t kRansL = 1ull << 31;

struct Sym { uint16_t start; uint16_t range; uint8_t bypass; };

inline void enc_put(uint64_t& x, uint32_t*& ptr, uint32_t start, uint32_t freq, int scale_bits) {
  const uint64_t x_max = ((kRansL >> scale_bits) << 32) * freq;
  if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
  x = ((x / freq) << scale_bits) + (x % freq) + start;
}
inline void enc_put_bits(uint64_t& x, uint32_t*& ptr, uint32_t val, int nbits) {
  const uint64_t freq = 1ull << (16 - nbits);
  const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
  if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
  x = (x << nbits) | val;
}
}  // namespace

extern "C" int lvae_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int32_t* cdf_out) {
  if (!pmf || !cdf_out || n <= 0 || precision <= 0 || precision > 30) return LVAE_E_BADARG;
  std::vector<uint32_t> cdf(n + 1);
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) {
    if (!(pmf[i] >= 0.f) || !isfinite(pmf[i])) return LVAE_E_BADARG;
    cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
  }
  uint32_t total = 0;
  for (uint32_t c : cdf) total += c;
  if (total == 0) return LVAE_E_BADARG;
  for (auto& c : cdf) c = (uint32_t)((((uint64_t)1 << precision) * c) / total);
  for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      uint32_t best_freq = ~0u; int best_steal = -1;
      for (int j = 0; j < n; ++j) {
        const uint32_t freq = cdf[j + 1] - cdf[j];
        if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
      }
      if (best_steal < 0) return LVAE_E_BADARG;
      if (best_steal < i) { for (int j = best_steal + 1; j <= i; ++j) cdf[j]--; }
      else { for (int j = i + 1; j <= best_steal; ++j) cdf[j]++; }
    }
  }
  for (int i = 0; i <= n; ++i) cdf_out[i] = (int32_t)cdf[i];
  return 0;
}

extern "C" int64_t lvae_rans_bound(int64_t n) {
  // every symbol emits at most one 32-bit word per rANS step; a bypassed symbol adds at most
  // 1 + 1 + 8 nibble steps (32-bit raw value).  Steps only emit a word when the state overflows,
  // i.e. at most once per 16 bits of payload, so 4 bytes per step is a safe bound.
  return (n * 11 + 4) * 4;
}

namespace {
// ---- encoder tables: ryg_rans' reciprocal form of  x = ((x / freq) << 16) + (x % freq) + start  (Rans64EncSymbolInit /
// Rans64EncPutSymbol): q = mulhi(x, rcp_freq) >> rcp_shift is exactly floor(x / freq) for every reachable state, so the
// bytes are those of the division form -- without a 64-bit division per symbol.
struct EncSym { uint64_t x_max, rcp_freq; uint32_t bias, cmpl_freq, rcp_shift, pad; };

inline void enc_sym_init(EncSym& e, uint32_t start, uint32_t freq) {
  e.x_max = ((kRansL >> kPrecision) << 32) * freq;
  e.cmpl_freq = (1u << kPrecision) - freq;
  e.pad = 0;
  if (freq < 2) {
    e.rcp_freq = ~0ull; e.rcp_shift = 0; e.bias = start + (1u << kPrecision) - 1;
  } else {
    uint32_t shift = 0;
    while (freq > (1u << shift)) ++shift;
    const unsigned __int128 num = ((unsigned __int128)1 << (shift + 63)) + freq - 1;
    e.rcp_freq = (uint64_t)(num / freq);
    e.rcp_shift = shift - 1;
    e.bias = start;
  }
}
inline void enc_put_fast(uint64_t& x, uint32_t*& ptr, const EncSym& e) {
  if (x >= e.x_max) { *--ptr = (uint32_t)x; x >>= 32; }
  const uint64_t q = (uint64_t)(((unsigned __int128)x * e.rcp_freq) >> 64) >> e.rcp_shift;
  x = x + e.bias + q * e.cmpl_freq;
}

struct EncTables {
  std::vector<EncSym> sym;        // [n_cdf][stride]: entry v of row r codes the interval [cdf[r][v], cdf[r][v+1])
  int stride = 0;
  void build(const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int n_cdf) {
    stride = cdf_stride;
    sym.assign((size_t)n_cdf * cdf_stride, EncSym{});
    for (int r = 0; r < n_cdf; ++r) {
      const int32_t* c = cdf + (int64_t)r * cdf_stride;
      for (int v = 0; v + 1 < cdf_len[r] && v + 1 < cdf_stride; ++v)
        enc_sym_init(sym[(size_t)r * cdf_stride + v], (uint32_t)c[v], (uint32_t)(c[v + 1] - c[v]));
    }
  }
};

// One stream, coded back to front straight from sym / idx (rANS is last-in first-out): for an out-of-table symbol the
// forward order is [table symbol, bypass-count prefix, nibbles 0..k-1], so the reverse pass emits nibbles k-1..0, then the
// prefix reversed, then the table symbol.  Words are written downwards from the end of a scratch buffer.
struct EncState {
  const int32_t* sym; const int32_t* idx; int64_t i;      // next symbol to code is i - 1
  uint64_t x; uint32_t* ptr; uint32_t* end;
  std::vector<uint32_t> buf;
};

// Scratch sizing without a pass over the data: a table symbol carries at most 16 bits, i.e. n of them emit at most
// n / 2 + 2 words; the rest of the n + 16 words is slack for out-of-table symbols (at most 3 words each), re-checked --
// and grown -- in that rare branch only.
int enc_open(EncState& e, const int32_t* sym, const int32_t* idx, int64_t n) {
  const size_t need = (size_t)n + 16;
  if (e.buf.size() < need) e.buf.resize(need);
  e.end = e.buf.data() + e.buf.size();
  e.ptr = e.end;
  e.x = kRansL;
  e.sym = sym; e.idx = idx; e.i = n;
  return 0;
}

void enc_grow(EncState& e) {
  const size_t used = (size_t)(e.end - e.ptr);
  std::vector<uint32_t> bigger(e.buf.size() * 2 + 64);
  memcpy(bigger.data() + bigger.size() - used, e.ptr, used * 4);
  e.buf.swap(bigger);
  e.end = e.buf.data() + e.buf.size();
  e.ptr = e.end - used;
}

// codes symbol i - 1; false: bad table index
inline bool enc_step(EncState& e, const EncTables& T, const int32_t* cdf_len, const int32_t* offset, int n_cdf) {
  const int64_t i = --e.i;
  const int32_t ci = e.idx[i];
  if ((uint32_t)ci >= (uint32_t)n_cdf) { e.i = 0; return false; }
  const int32_t max_value = cdf_len[ci] - 2;
  int32_t value = e.sym[i] - offset[ci];
  if (value < 0 || value >= max_value) {
    // words still free must cover this symbol (<= 3) and the table symbols left (<= i / 2 + 2)
    if ((int64_t)(e.ptr - e.buf.data()) < i / 2 + 8) enc_grow(e);
    uint32_t raw_val;
    if (value < 0) raw_val = (uint32_t)(-2 * (int64_t)value - 1);
    else raw_val = (uint32_t)(2 * ((int64_t)value - max_value));
    value = max_value;
    int32_t n_bypass = 0;
    while (n_bypass < 8 && (raw_val >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
    for (int32_t j = n_bypass; j-- > 0;) enc_put_bits(e.x, e.ptr, (raw_val >> (j * kBypassPrecision)) & kMaxBypassVal, kBypassPrecision);
    enc_put_bits(e.x, e.ptr, (uint32_t)(n_bypass % kMaxBypassVal), kBypassPrecision);
    for (int32_t m = n_bypass / kMaxBypassVal; m-- > 0;) enc_put_bits(e.x, e.ptr, kMaxBypassVal, kBypassPrecision);
  }
  enc_put_fast(e.x, e.ptr, T.sym[(size_t)ci * T.stride + value]);
  return true;
}

int enc_close(EncState& e, uint8_t* out, int64_t out_cap, int64_t* out_len) {
  e.ptr -= 2; e.ptr[0] = (uint32_t)e.x; e.ptr[1] = (uint32_t)(e.x >> 32);
  const int64_t nbytes = (int64_t)(e.end - e.ptr) * 4;
  if (nbytes > out_cap) return LVAE_E_BADARG;
  memcpy(out, e.ptr, (size_t)nbytes);
  *out_len = nbytes;
  return 0;
}

int encode_stream(const int32_t* sym, const int32_t* idx, int64_t n, const EncTables& T, const int32_t* cdf_len,
                  const int32_t* offset, int n_cdf, uint8_t* out, int64_t out_cap, int64_t* out_len) {
  static thread_local EncState e;
  enc_open(e, sym, idx, n);
  bool ok = true;
  while (e.i > 0) ok &= enc_step(e, T, cdf_len, offset, n_cdf);
  if (!ok) return LVAE_E_BADARG;
  return enc_close(e, out, out_cap, out_len);
}

// ---- decoder tables: per row, the symbol that contains cumulative value b << 8 (256 buckets): the search for
// c[s] <= cum < c[s+1] starts there and walks up, one or two comparisons for almost every symbol
struct DecTables {
  std::vector<uint16_t> lut;      // [n_cdf][256]
  void build(const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int n_cdf) {
    lut.assign((size_t)n_cdf * 256, 0);
    for (int r = 0; r < n_cdf; ++r) {
      const int32_t* c = cdf + (int64_t)r * cdf_stride;
      const int last = cdf_len[r] - 2;            // last symbol of the row
      int s = 0;
      for (int b = 0; b < 256; ++b) {
        const uint32_t cum = (uint32_t)b << 8;
        while (s < last && (uint32_t)c[s + 1] <= cum) ++s;
        lut[(size_t)r * 256 + b] = (uint16_t)s;
      }
    }
  }
};

struct DecState {
  const uint32_t* words; int64_t nwords, pos; uint64_t x;
  const int32_t* idx; int32_t* out; int64_t n, i; int rc;
};

inline int dec_open(DecState& d, const uint8_t* in, int64_t in_len, const int32_t* idx, int64_t n, int32_t* out) {
  if (in_len < 8 || (in_len & 3) || ((uintptr_t)in & 3)) return LVAE_E_CORRUPT;
  d.words = reinterpret_cast<const uint32_t*>(in); d.nwords = in_len / 4; d.pos = 2;
  d.x = (uint64_t)d.words[0] | ((uint64_t)d.words[1] << 32);
  d.idx = idx; d.out = out; d.n = n; d.i = 0; d.rc = 0;
  return 0;
}

// decode symbol d.i of the stream; returns false when the stream has ended (d.rc tells how)
inline bool dec_step(DecState& d, const DecTables& T, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                     const int32_t* offset, int n_cdf) {
  if (d.i >= d.n) return false;
  const int32_t ci = d.idx[d.i];
  if (ci < 0 || ci >= n_cdf) { d.rc = LVAE_E_BADARG; d.i = d.n; return false; }
  const int32_t* c = cdf + (int64_t)ci * cdf_stride;
  const int32_t max_value = cdf_len[ci] - 2;
  const uint32_t cum = (uint32_t)(d.x & 0xffffu);
  int s = T.lut[(size_t)ci * 256 + (cum >> 8)];
  while ((uint32_t)c[s + 1] <= cum) ++s;                 // rows end at 65536 > cum: the walk stops inside the row
  d.x = (uint64_t)(uint32_t)(c[s + 1] - c[s]) * (d.x >> kPrecision) + cum - (uint32_t)c[s];
  bool overrun = false;
  auto renorm = [&]() {
    if (d.x < kRansL) {
      if (d.pos >= d.nwords) { overrun = true; return; }
      d.x = (d.x << 32) | d.words[d.pos++];
    }
  };
  renorm();
  int32_t value = s;
  if (value == max_value) {
    auto get_bits = [&]() -> int32_t {
      const int32_t v = (int32_t)(d.x & kMaxBypassVal);
      d.x >>= kBypassPrecision;
      renorm();
      return v;
    };
    int32_t val = get_bits();
    int32_t n_bypass = val;
    while (val == kMaxBypassVal && !overrun) { val = get_bits(); n_bypass += val; }
    if (n_bypass > 8) { d.rc = LVAE_E_CORRUPT; d.i = d.n; return false; }
    uint32_t raw_val = 0;
    for (int32_t j = 0; j < n_bypass; ++j) raw_val |= (uint32_t)get_bits() << (j * kBypassPrecision);
    value = (int32_t)(raw_val >> 1);
    if (raw_val & 1) value = -value - 1; else value += max_value;
  }
  if (overrun) { d.rc = LVAE_E_CORRUPT; d.i = d.n; return false; }
  d.out[d.i++] = value + offset[ci];
  return true;
}

bool tables_ok(const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int n_cdf) {
  if (n_cdf <= 0 || cdf_stride < 2) return false;
  for (int r = 0; r < n_cdf; ++r) {
    if (cdf_len[r] < 2 || cdf_len[r] > cdf_stride) return false;
    const int32_t* c = cdf + (int64_t)r * cdf_stride;
    if (c[0] != 0 || c[cdf_len[r] - 1] != (1 << kPrecision)) return false;
    for (int v = 0; v + 1 < cdf_len[r]; ++v) if (c[v + 1] <= c[v]) return false;
  }
  return true;
}
}  // namespace

extern "C" int lvae_rans_encode(const int32_t* sym, const int32_t* idx, int64_t n,
                                const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                const int32_t* offset, int n_cdf, uint8_t* out, int64_t out_cap,
                                int64_t* out_len) {
  if (!sym || !idx || !cdf || !cdf_len || !offset || !out || !out_len || n < 0) return LVAE_E_BADARG;
  if (!tables_ok(cdf, cdf_stride, cdf_len, n_cdf)) return LVAE_E_BADARG;
  EncTables T;
  T.build(cdf, cdf_stride, cdf_len, n_cdf);
  return encode_stream(sym, idx, n, T, cdf_len, offset, n_cdf, out, out_cap, out_len);
}

extern "C" int lvae_rans_decode(const uint8_t* in, int64_t in_len, const int32_t* idx, int64_t n,
                                const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                const int32_t* offset, int n_cdf, int32_t* sym_out) {
  if (!in || !idx || !cdf || !cdf_len || !offset || !sym_out || n < 0) return LVAE_E_BADARG;
  if (!tables_ok(cdf, cdf_stride, cdf_len, n_cdf)) return LVAE_E_BADARG;
  DecTables T;
  T.build(cdf, cdf_stride, cdf_len, n_cdf);
  std::vector<uint32_t> aligned;
  if ((uintptr_t)in & 3) {                              // the word reads want 4-byte alignment
    if (in_len < 8 || (in_len & 3)) return LVAE_E_CORRUPT;
    aligned.resize((size_t)in_len / 4);
    memcpy(aligned.data(), in, (size_t)in_len);
    in = reinterpret_cast<const uint8_t*>(aligned.data());
  }
  DecState d;
  int rc = dec_open(d, in, in_len, idx, n, sym_out);
  if (rc) return rc;
  while (dec_step(d, T, cdf, cdf_stride, cdf_len, offset, n_cdf)) {}
  return d.rc;
}

// Encode `n_streams` independent streams (one per (image, layer)) on up to `n_threads` host threads.  Stream i
// covers symbols [begin[i], begin[i+1]) of sym / idx and uses table set table_of[i] (cdf_sets[k] etc. describe table
// set k; all streams of a model usually share one).  Its bytes are written at out + out_begin[i] (capacity
// out_begin[i+1] - out_begin[i] >= lvae_rans_bound(count)), their count to out_len[i].
extern "C" int lvae_rans_encode_streams(const int32_t* sym, const int32_t* idx, const int64_t* begin, int n_streams,
                                        const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                        const int32_t* offset, int n_cdf, uint8_t* out, const int64_t* out_begin,
                                        int64_t* out_len, int n_threads) {
  if (!sym || !idx || !begin || !out || !out_begin || !out_len || n_streams < 0) return LVAE_E_BADARG;
  if (!cdf || !cdf_len || !offset || !tables_ok(cdf, cdf_stride, cdf_len, n_cdf)) return LVAE_E_BADARG;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_streams) n_threads = n_streams;
  EncTables T;                                          // built once, shared by all streams of the call
  T.build(cdf, cdf_stride, cdf_len, n_cdf);
  std::atomic<int> next(0), status(0);
  const int pair = n_streams >= 2 * n_threads ? 2 : 1;   // two streams per worker in lockstep (see the decoder)
  auto work = [&]() {
    static thread_local EncState e[2];
    for (;;) {
      const int i0 = next.fetch_add(pair);
      if (i0 >= n_streams) return;
      int which[2], live = 0;
      for (int k = 0; k < pair && i0 + k < n_streams; ++k) {
        const int i = i0 + k;
        enc_open(e[live], sym + begin[i], idx + begin[i], begin[i + 1] - begin[i]);
        which[live++] = i;
      }
      bool ok = true;
      if (live == 2) {
        while (e[0].i > 0 && e[1].i > 0) {
          ok &= enc_step(e[0], T, cdf_len, offset, n_cdf);
          ok &= enc_step(e[1], T, cdf_len, offset, n_cdf);
        }
      }
      for (int k = 0; k < live; ++k) {
        while (e[k].i > 0) ok &= enc_step(e[k], T, cdf_len, offset, n_cdf);
        const int i = which[k];
        const int rc = ok ? enc_close(e[k], out + out_begin[i], out_begin[i + 1] - out_begin[i], out_len + i) : LVAE_E_BADARG;
        if (rc != 0) status.store(rc);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return status.load();
}

// Decode `n_streams` independent streams on up to `n_threads` host threads: stream i reads bytes
// [in_begin[i], in_begin[i+1]) of `in` and produces symbols [begin[i], begin[i+1]) of sym_out from the same range of idx.
// A worker takes two streams at a time and advances them in lockstep: the decoder is one serial dependency chain per
// stream (state -> table search -> multiply -> renormalise), two chains in flight hide each other's latencies.
extern "C" int lvae_rans_decode_streams(const uint8_t* in, const int64_t* in_begin, const int32_t* idx, const int64_t* begin,
                                        int n_streams, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                        const int32_t* offset, int n_cdf, int32_t* sym_out, int n_threads) {
  if (!in || !in_begin || !idx || !begin || !sym_out || n_streams < 0) return LVAE_E_BADARG;
  if (!cdf || !cdf_len || !offset || !tables_ok(cdf, cdf_stride, cdf_len, n_cdf)) return LVAE_E_BADARG;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_streams) n_threads = n_streams;
  DecTables T;
  T.build(cdf, cdf_stride, cdf_len, n_cdf);
  std::atomic<int> next(0), status(0);
  const int pair = n_streams >= 2 * n_threads ? 2 : 1;   // pair streams only when every thread still gets work
  auto work = [&]() {
    std::vector<uint32_t> copy[2];
    for (;;) {
      const int i0 = next.fetch_add(pair);
      if (i0 >= n_streams) return;
      DecState d[2];
      int live = 0;
      for (int k = 0; k < pair && i0 + k < n_streams; ++k) {
        const int i = i0 + k;
        const uint8_t* src = in + in_begin[i];
        const int64_t len = in_begin[i + 1] - in_begin[i];
        if (((uintptr_t)src & 3) && len >= 8 && !(len & 3)) {
          copy[k].resize((size_t)len / 4);
          memcpy(copy[k].data(), src, (size_t)len);
          src = reinterpret_cast<const uint8_t*>(copy[k].data());
        }
        const int rc = dec_open(d[live], src, len, idx + begin[i], begin[i + 1] - begin[i], sym_out + begin[i]);
        if (rc != 0) { status.store(rc); continue; }
        ++live;
      }
      if (live == 2) {
        while (d[0].i < d[0].n && d[1].i < d[1].n) {
          dec_step(d[0], T, cdf, cdf_stride, cdf_len, offset, n_cdf);
          dec_step(d[1], T, cdf, cdf_stride, cdf_len, offset, n_cdf);
        }
      }
      for (int k = 0; k < live; ++k) {
        while (dec_step(d[k], T, cdf, cdf_stride, cdf_len, offset, n_cdf)) {}
        if (d[k].rc != 0) status.store(d[k].rc);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return status.load();
}
