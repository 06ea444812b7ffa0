// Host entropy coder: quantised-CDF builder and rANS64 encoder/decoder.
//
// Restates the published behaviour of CompressAI (un-vendored dependency of the reference; call
// sites lvae/models/qarv/model.py:106-113,123-124): cpp_exts/ops/ops.cpp pmf_to_quantized_cdf,
// cpp_exts/rans/rans_interface.cpp {encode,decode}_with_indexes and ryg_rans' rans64.h -- 64-bit
// state, 32-bit renormalisation words, 16-bit probabilities, 4-bit bypass for out-of-table symbols.
// Byte-compatibility with a real CompressAI build is "parity unpinned" (absent offline); the
// testable properties are equality with the Python restatement in oracle/ and lossless round trips.
// The coder is serial per stream and thread-safe: callers run one stream per (image, layer).
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

extern "C" int lvae_rans_encode(const int32_t* sym, const int32_t* idx, int64_t n,
                                const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                const int32_t* offset, int n_cdf, uint8_t* out, int64_t out_cap,
                                int64_t* out_len) {
  if (!sym || !idx || !cdf || !cdf_len || !offset || !out || !out_len || n < 0) return LVAE_E_BADARG;
  std::vector<Sym> syms;
  syms.reserve((size_t)n + 16);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = idx[i];
    if (ci < 0 || ci >= n_cdf) return LVAE_E_BADARG;
    const int32_t* c = cdf + (int64_t)ci * cdf_stride;
    const int32_t max_value = cdf_len[ci] - 2;
    int32_t value = sym[i] - offset[ci];
    uint32_t raw_val = 0;
    if (value < 0) { raw_val = (uint32_t)(-2 * (int64_t)value - 1); value = max_value; }
    else if (value >= max_value) { raw_val = (uint32_t)(2 * ((int64_t)value - max_value)); value = max_value; }
    syms.push_back({(uint16_t)c[value], (uint16_t)(c[value + 1] - c[value]), 0});
    if (value == max_value) {
      int32_t n_bypass = 0;
      while (n_bypass < 8 && (raw_val >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
      int32_t val = n_bypass;
      while (val >= kMaxBypassVal) { syms.push_back({(uint16_t)kMaxBypassVal, 0, 1}); val -= kMaxBypassVal; }
      syms.push_back({(uint16_t)val, 0, 1});
      for (int32_t j = 0; j < n_bypass; ++j)
        syms.push_back({(uint16_t)((raw_val >> (j * kBypassPrecision)) & kMaxBypassVal), 0, 1});
    }
  }
  std::vector<uint32_t> buf(syms.size() + 2);
  uint32_t* end = buf.data() + buf.size();
  uint32_t* ptr = end;
  uint64_t x = kRansL;
  for (size_t i = syms.size(); i-- > 0;) {
    const Sym& s = syms[i];
    if (!s.bypass) enc_put(x, ptr, s.start, s.range, kPrecision);
    else enc_put_bits(x, ptr, s.start, kBypassPrecision);
  }
  ptr -= 2; ptr[0] = (uint32_t)x; ptr[1] = (uint32_t)(x >> 32);
  const int64_t nbytes = (int64_t)(end - ptr) * 4;
  if (nbytes > out_cap) return LVAE_E_BADARG;
  memcpy(out, ptr, (size_t)nbytes);
  *out_len = nbytes;
  return 0;
}

extern "C" int lvae_rans_decode(const uint8_t* in, int64_t in_len, const int32_t* idx, int64_t n,
                                const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                const int32_t* offset, int n_cdf, int32_t* sym_out) {
  if (!in || !idx || !cdf || !cdf_len || !offset || !sym_out || n < 0) return LVAE_E_BADARG;
  if (in_len < 8 || (in_len & 3)) return LVAE_E_CORRUPT;
  const int64_t nwords = in_len / 4;
  std::vector<uint32_t> words((size_t)nwords);
  memcpy(words.data(), in, (size_t)in_len);
  int64_t pos = 2;
  uint64_t x = (uint64_t)words[0] | ((uint64_t)words[1] << 32);
  const uint64_t mask = (1ull << kPrecision) - 1;
  bool overrun = false;
  auto renorm = [&]() {
    if (x < kRansL) {
      if (pos >= nwords) { overrun = true; return; }
      x = (x << 32) | words[pos++];
    }
  };
  auto get_bits = [&]() -> int32_t {
    const int32_t v = (int32_t)(x & kMaxBypassVal);
    x >>= kBypassPrecision;
    renorm();
    return v;
  };
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = idx[i];
    if (ci < 0 || ci >= n_cdf) return LVAE_E_BADARG;
    const int32_t* c = cdf + (int64_t)ci * cdf_stride;
    const int32_t size = cdf_len[ci];
    const int32_t max_value = size - 2;
    const uint32_t cum = (uint32_t)(x & mask);
    // the entry s with c[s] <= cum < c[s+1] (rows are strictly increasing).  The tables are discretised Gaussians
    // centred on the row: walk outwards from the mode (1-3 comparisons for almost every symbol), falling back to a
    // binary search once the walk has left the bulk
    int s = -offset[ci];                 // centre of the row (offset = -ceil(6.1 sigma))
    if (s < 0 || s > size - 2) s = (size - 2) >> 1;
    if ((uint32_t)c[s] <= cum) {
      int steps = 0;
      while ((uint32_t)c[s + 1] <= cum) {
        if (++steps > 4) { int lo = s + 1, hi = size - 1; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((uint32_t)c[mid] <= cum) lo = mid; else hi = mid; } s = lo; break; }
        ++s;
      }
    } else {
      int steps = 0;
      do {
        if (++steps > 4) { int lo = 0, hi = s; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((uint32_t)c[mid] <= cum) lo = mid; else hi = mid; } s = lo; break; }
        --s;
      } while ((uint32_t)c[s] > cum);
    }
    x = (uint64_t)(c[s + 1] - c[s]) * (x >> kPrecision) + cum - (uint32_t)c[s];
    renorm();
    int32_t value = s;
    if (value == max_value) {
      int32_t val = get_bits();
      int32_t n_bypass = val;
      while (val == kMaxBypassVal && !overrun) { val = get_bits(); n_bypass += val; }
      if (n_bypass > 8) return LVAE_E_CORRUPT;
      uint32_t raw_val = 0;
      for (int32_t j = 0; j < n_bypass; ++j) raw_val |= (uint32_t)get_bits() << (j * kBypassPrecision);
      value = (int32_t)(raw_val >> 1);
      if (raw_val & 1) value = -value - 1; else value += max_value;
    }
    if (overrun) return LVAE_E_CORRUPT;
    sym_out[i] = value + offset[ci];
  }
  return 0;
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
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_streams) n_threads = n_streams;
  std::atomic<int> next(0), status(0);
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_streams) return;
      const int rc = lvae_rans_encode(sym + begin[i], idx + begin[i], begin[i + 1] - begin[i], cdf, cdf_stride, cdf_len,
                                      offset, n_cdf, out + out_begin[i], out_begin[i + 1] - out_begin[i], out_len + i);
      if (rc != 0) status.store(rc);
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
extern "C" int lvae_rans_decode_streams(const uint8_t* in, const int64_t* in_begin, const int32_t* idx, const int64_t* begin,
                                        int n_streams, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                                        const int32_t* offset, int n_cdf, int32_t* sym_out, int n_threads) {
  if (!in || !in_begin || !idx || !begin || !sym_out || n_streams < 0) return LVAE_E_BADARG;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_streams) n_threads = n_streams;
  std::atomic<int> next(0), status(0);
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_streams) return;
      const int rc = lvae_rans_decode(in + in_begin[i], in_begin[i + 1] - in_begin[i], idx + begin[i], begin[i + 1] - begin[i],
                                      cdf, cdf_stride, cdf_len, offset, n_cdf, sym_out + begin[i]);
      if (rc != 0) status.store(rc);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return status.load();
}
