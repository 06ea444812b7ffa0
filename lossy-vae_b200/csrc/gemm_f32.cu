// fp32 conv-as-GEMM on CUDA cores (LVAE_PREC_FP32): the arithmetic closest to the CPU fp32
// reference (true fp32 products, fp32 accumulation).  One kernel covers 1x1 / 3x3 / patch (k=s)
// convolutions over NHWC activations with an optional second K segment (the post_merge concat),
// and the fused epilogues of lvae_b200.h.  The tensor-core path lives in gemm_tc.cu.
//
// Tiling: BM=128 x BN in {128,64,32} x BK=16, 256 threads, each thread a (2x4) x (TNx) register
// tile split in two 4-wide groups so that shared-memory reads are conflict-free 128-bit loads.
// K order is fixed and there is no split-K: results are bit-reproducible and batch-invariant
// (SURVEY F12).
#include "common.cuh"

namespace lvae {

struct GemmParams {
  const float* a0; const float* a1;
  int B, H, W, Ho, Wo, C0, C1, ks, stride, pad;
  const float* w; const float* bias; int N, K, M;
  int epi; const float* gamma; const float* res; float* out; int r;
  int a_act;     // 1: GELU on every A element as it is read (VDBlock c_i(gelu(x)))
};

constexpr int BM = 128, BK = 16, NT = 256;

__device__ __forceinline__ void store_out(const GemmParams& p, int m, int n, float v) {
  if (p.epi == LVAE_EPI_SHUFFLE_NHWC || p.epi == LVAE_EPI_SHUFFLE_NCHW) {
    const int r = p.r, Co = p.N / (r * r);
    const int q = n / Co, c = n - q * Co, i = q / r, j = q - i * r;
    const int wo = m % p.Wo; const int t = m / p.Wo; const int ho = t % p.Ho; const int b = t / p.Ho;
    const int Hr = p.Ho * r, Wr = p.Wo * r;
    if (p.epi == LVAE_EPI_SHUFFLE_NHWC)
      p.out[(((int64_t)b * Hr + ho * r + i) * Wr + wo * r + j) * Co + c] = v;
    else
      p.out[(((int64_t)b * Co + c) * Hr + ho * r + i) * Wr + wo * r + j] = v;
  } else {
    p.out[(int64_t)m * p.N + n] = v;
  }
}

__device__ __forceinline__ float apply_epi(const GemmParams& p, int m, int n, float acc) {
  float v = acc;
  if (p.bias) v = __fadd_rn(v, p.bias[n]);
  switch (p.epi) {
    case LVAE_EPI_BIAS_GELU: v = gelu_erf(v); break;
    case LVAE_EPI_SCALE_RES: v = __fadd_rn(__fmul_rn(v, p.gamma[n]), p.res[(int64_t)m * p.N + n]); break;
    case LVAE_EPI_BIAS_RES:  v = __fadd_rn(p.res[(int64_t)m * p.N + n], v); break;
    default: break;
  }
  return v;
}

template <int BN>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(const GemmParams p) {
  constexpr int TN = BN / 16;          // columns per thread (8, 4, 2)
  constexpr int TNH = TN / 2;          // per 2 groups
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A loader: thread loads 2 float4: rows (tid>>2) and (tid>>2)+64, k quad (tid&3) ----
  const int a_kq = tid & 3;
  int a_row[2]; int a_b[2], a_h0[2], a_w0[2]; bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    a_row[i] = (tid >> 2) + 64 * i;
    const int m = m0 + a_row[i];
    a_ok[i] = m < p.M;
    const int mm = a_ok[i] ? m : 0;
    const int wo = mm % p.Wo; const int t = mm / p.Wo; const int ho = t % p.Ho;
    a_b[i] = t / p.Ho; a_h0[i] = ho * p.stride - p.pad; a_w0[i] = wo * p.stride - p.pad;
  }
  // ---- B loader: BN rows x 16 k = BN*4 float4; thread loads BN/64 of them ----
  constexpr int BLD = BN * 4 / NT > 0 ? BN * 4 / NT : 1;
  const int K0 = p.ks * p.ks * p.C0;

  float4 ra[2]; float4 rb[BLD];
  auto load_tiles = [&](int k0) {
    const int k = k0 + a_kq * 4;
    // decode k -> (segment, tap, c)
    const float* base = nullptr; int C = 0, c = 0, ky = 0, kx = 0; bool kvalid = k < p.K;
    if (kvalid) {
      if (k < K0) { const int tap = k / p.C0; c = k - tap * p.C0; ky = tap / p.ks; kx = tap - ky * p.ks; base = p.a0; C = p.C0; }
      else { c = k - K0; base = p.a1; C = p.C1; }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kvalid && a_ok[i]) {
        const int hh = a_h0[i] + ky, ww = a_w0[i] + kx;
        if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
          v = __ldg(reinterpret_cast<const float4*>(base + (((int64_t)a_b[i] * p.H + hh) * p.W + ww) * C + c));
      }
      if (p.a_act) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < BLD; ++i) {
      const int f = tid + i * NT;            // float4 index within the tile
      const int row = f >> 2, kq = f & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (BN * 4 >= NT || f < BN * 4) {
        const int n = n0 + row, kk = k0 + kq * 4;
        if (n < p.N && kk < p.K) v = __ldg(reinterpret_cast<const float4*>(p.w + (int64_t)n * p.K + kk));
      }
      rb[i] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][a_kq * 4 + 0][a_row[i]] = ra[i].x; As[buf][a_kq * 4 + 1][a_row[i]] = ra[i].y;
      As[buf][a_kq * 4 + 2][a_row[i]] = ra[i].z; As[buf][a_kq * 4 + 3][a_row[i]] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < BLD; ++i) {
      const int f = tid + i * NT;
      if (BN * 4 >= NT || f < BN * 4) {
        const int row = f >> 2, kq = f & 3;
        Bs[buf][kq * 4 + 0][row] = rb[i].x; Bs[buf][kq * 4 + 1][row] = rb[i].y;
        Bs[buf][kq * 4 + 2][row] = rb[i].z; Bs[buf][kq * 4 + 3][row] = rb[i].w;
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[TN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      if constexpr (TNH == 4) {
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
      } else if constexpr (TNH == 2) {
        *reinterpret_cast<float2*>(&b[0]) = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        *reinterpret_cast<float2*>(&b[2]) = *reinterpret_cast<const float2*>(&Bs[buf][k][BN / 2 + tx * 2]);
      } else {
        b[0] = Bs[buf][k][tx]; b[1] = Bs[buf][k][BN / 2 + tx];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) { store_tiles(buf ^ 1); }
    __syncthreads();
  }

  // ---- epilogue ----
  const bool plain = !(p.epi == LVAE_EPI_SHUFFLE_NHWC || p.epi == LVAE_EPI_SHUFFLE_NCHW);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int nb = n0 + g * (BN / 2) + tx * TNH;
      float v[TNH];
#pragma unroll
      for (int j = 0; j < TNH; ++j) v[j] = (nb + j < p.N) ? apply_epi(p, m, nb + j, acc[i][g * TNH + j]) : 0.f;
      if (plain && TNH == 4 && (p.N & 3) == 0 && nb + 3 < p.N) {
        *reinterpret_cast<float4*>(p.out + (int64_t)m * p.N + nb) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < TNH; ++j) if (nb + j < p.N) store_out(p, m, nb + j, v[j]);
      }
    }
  }
}

// Rank-K update for tiny K (z_proj: feature += W z + b with K = zdim in {8, 32, 96}, qarv/model.py:72-75):
// bandwidth-bound, so plain fp32 FMAs; the [N, K] weight tile of a CTA sits in shared memory, each thread owns
// one output column for a strip of rows and streams the residual in and the result out with coalesced accesses.
template <int KMAX>
__global__ void __launch_bounds__(256) gemm_smallk_kernel(const GemmParams p) {
  extern __shared__ float sk_smem[];                   // [256][K + 1] weights | [32][K] A rows
  const int K = p.K, n0 = blockIdx.y * 256, tid = threadIdx.x;
  float* ws = sk_smem; float* as = sk_smem + 256 * (K + 1);
  for (int i = tid; i < 256 * K; i += 256) {
    const int n = i / K, k = i - n * K;
    ws[n * (K + 1) + k] = (n0 + n < p.N) ? __ldg(p.w + (int64_t)(n0 + n) * K + k) : 0.f;
  }
  const int n = n0 + tid;
  const bool n_ok = n < p.N;
  const float b = (n_ok && p.bias) ? __ldg(p.bias + n) : 0.f;
  const float g = (n_ok && p.epi == LVAE_EPI_SCALE_RES) ? __ldg(p.gamma + n) : 1.f;
  const float* wrow = ws + tid * (K + 1);
  for (int m0 = blockIdx.x * 32; m0 < p.M; m0 += gridDim.x * 32) {
    __syncthreads();
    const int rows = (p.M - m0) < 32 ? (p.M - m0) : 32;
    for (int i = tid; i < rows * K; i += 256) as[i] = __ldg(p.a0 + (int64_t)m0 * K + i);
    __syncthreads();
    if (!n_ok) continue;
    float rr[32];
    const bool has_res = (p.epi == LVAE_EPI_SCALE_RES || p.epi == LVAE_EPI_BIAS_RES);
    if (has_res) {
#pragma unroll
      for (int r = 0; r < 32; ++r) rr[r] = (r < rows) ? p.res[(int64_t)(m0 + r) * p.N + n] : 0.f;
    }
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      if (r >= rows) break;
      float acc = 0.f;
      const float* arow = as + r * K;
#pragma unroll 8
      for (int k = 0; k < K; ++k) acc = fmaf(arow[k], wrow[k], acc);
      float v = __fadd_rn(acc, b);
      if (p.epi == LVAE_EPI_BIAS_GELU) v = gelu_erf(v);
      else if (p.epi == LVAE_EPI_SCALE_RES) v = __fadd_rn(__fmul_rn(v, g), rr[r]);
      else if (p.epi == LVAE_EPI_BIAS_RES) v = __fadd_rn(rr[r], v);
      p.out[(int64_t)(m0 + r) * p.N + n] = v;
    }
  }
}

// Streaming form for N % 4 == 0 (every z_proj of the registered models): the weights sit transposed in shared memory
// ([K][N], so a thread reads 4 consecutive columns of one k with a conflict-free 128-bit load), a block owns strips of
// SK2_ROWS rows whose A values are staged in shared memory (broadcast reads), and every thread produces float4 pieces
// of the output: residual in / result out with 128-bit coalesced accesses, 1.25 K shared-memory instructions per 4
// outputs.  The per-output FMA chain runs over ascending k exactly as in gemm_smallk_kernel (bit-identical).  That
// kernel spent 254 M warp instructions on 56 M FFMAs (64-bit index arithmetic, a spilled residual array, two
// __syncthreads per 32 rows) and ran qres34m's K = 24 update at 1.1 TB/s.
constexpr int SK2_ROWS = 64;
__global__ void __launch_bounds__(256) gemm_smallk2_kernel(const GemmParams p) {
  extern __shared__ __align__(16) float sk2_smem[];    // wT [K][N] | as [SK2_ROWS][K]
  const int K = p.K, N = p.N, Q = N >> 2, tid = threadIdx.x;
  float* wT = sk2_smem;
  float* as = sk2_smem + K * N;
  for (int i = tid; i < K * N; i += 256) {
    const int n = i / K, k = i - n * K;                  // coalesced global read, transposed shared write
    wT[k * N + n] = __ldg(p.w + i);
  }
  const bool scale_res = p.epi == LVAE_EPI_SCALE_RES;
  const bool has_res = scale_res || p.epi == LVAE_EPI_BIAS_RES;
  for (int m0 = blockIdx.x * SK2_ROWS; m0 < p.M; m0 += gridDim.x * SK2_ROWS) {
    __syncthreads();
    const int rows = (p.M - m0) < SK2_ROWS ? (p.M - m0) : SK2_ROWS;
    for (int i = tid; i < rows * K; i += 256) as[i] = __ldg(p.a0 + (int64_t)m0 * K + i);
    __syncthreads();
    const float* res0 = has_res ? p.res + (int64_t)m0 * N : nullptr;
    float* out0 = p.out + (int64_t)m0 * N;
#pragma unroll 2
    for (int i = tid; i < rows * Q; i += 256) {
      const int r = i / Q, n = (i - r * Q) << 2;
      const int o = r * N + n;
      float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_res) rr = *reinterpret_cast<const float4*>(res0 + o);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* ar = as + r * K;
      for (int k = 0; k < K; k += 4) {                   // K % 4 == 0
        const float4 a = *reinterpret_cast<const float4*>(ar + k);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 w4 = *reinterpret_cast<const float4*>(wT + (k + kk) * N + n);
          acc.x = fmaf(av[kk], w4.x, acc.x); acc.y = fmaf(av[kk], w4.y, acc.y);
          acc.z = fmaf(av[kk], w4.z, acc.z); acc.w = fmaf(av[kk], w4.w, acc.w);
        }
      }
      const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v = make_float4(__fadd_rn(acc.x, b4.x), __fadd_rn(acc.y, b4.y), __fadd_rn(acc.z, b4.z), __fadd_rn(acc.w, b4.w));
      if (p.epi == LVAE_EPI_BIAS_GELU) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
      } else if (scale_res) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
        v.x = __fadd_rn(__fmul_rn(v.x, g4.x), rr.x); v.y = __fadd_rn(__fmul_rn(v.y, g4.y), rr.y);
        v.z = __fadd_rn(__fmul_rn(v.z, g4.z), rr.z); v.w = __fadd_rn(__fmul_rn(v.w, g4.w), rr.w);
      } else if (has_res) {
        v.x = __fadd_rn(rr.x, v.x); v.y = __fadd_rn(rr.y, v.y); v.z = __fadd_rn(rr.z, v.z); v.w = __fadd_rn(rr.w, v.w);
      }
      *reinterpret_cast<float4*>(out0 + o) = v;
    }
  }
}

bool gemm_smallk_applicable(const lvae_gemm_desc* d) {
  const int K = d->ksize * d->ksize * d->C0 + (d->a1 ? d->C1 : 0);
  return d->ksize == 1 && d->stride == 1 && d->pad == 0 && d->a1 == nullptr && d->a0 != nullptr && d->out != nullptr &&
         K <= 32 && d->N >= 128 && d->out_planes[0] == nullptr && d->a_act == 0 &&
         (d->epilogue == LVAE_EPI_BIAS || d->epilogue == LVAE_EPI_BIAS_RES || d->epilogue == LVAE_EPI_SCALE_RES);
}

int gemm_smallk_launch(const lvae_gemm_desc* d, cudaStream_t stream) {
  GemmParams p;
  p.a0 = d->a0; p.a1 = nullptr; p.B = d->B; p.H = d->H; p.W = d->W; p.Ho = d->H; p.Wo = d->W;
  p.C0 = d->C0; p.C1 = 0; p.ks = 1; p.stride = 1; p.pad = 0;
  p.w = d->w; p.bias = d->bias; p.N = d->N; p.K = d->C0; p.M = d->B * d->H * d->W;
  p.epi = d->epilogue; p.gamma = d->gamma; p.res = d->res; p.out = d->out; p.r = 0; p.a_act = 0;
  if (p.M == 0) return 0;
  const int smem = (256 * (p.K + 1) + 32 * p.K) * 4;
  static bool configured = false;
  if (!configured) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(gemm_smallk_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (256 * 97 + 32 * 96) * 4));
    configured = true;
  }
  const int ny = (p.N + 255) / 256;
  int nx = (p.M + 31) / 32;
  const int cap = 148 * 4 / ny > 0 ? 148 * 4 / ny : 1;
  if (nx > cap) nx = cap;
  if (p.N % 4 == 0 && p.K % 4 == 0) {
    const int smem2 = (p.K * p.N + SK2_ROWS * p.K) * 4;
    static int configured2 = 0;
    if (smem2 > 48 * 1024 && smem2 > configured2) {
      LVAE_CUDA_CALL(cudaFuncSetAttribute(gemm_smallk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      configured2 = smem2;
    }
    int nx2 = (p.M + SK2_ROWS - 1) / SK2_ROWS;
    if (nx2 > 148 * 4) nx2 = 148 * 4;
    gemm_smallk2_kernel<<<nx2, 256, smem2, stream>>>(p);
  } else {
    gemm_smallk_kernel<96><<<dim3(nx, ny), 256, smem, stream>>>(p);
  }
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

int gemm_f32_launch(const lvae_gemm_desc* d, cudaStream_t stream) {
  GemmParams p;
  p.a0 = d->a0; p.a1 = d->a1; p.B = d->B; p.H = d->H; p.W = d->W;
  p.C0 = d->C0; p.C1 = d->a1 ? d->C1 : 0; p.ks = d->ksize; p.stride = d->stride; p.pad = d->pad;
  p.Ho = (d->H + 2 * d->pad - d->ksize) / d->stride + 1;
  p.Wo = (d->W + 2 * d->pad - d->ksize) / d->stride + 1;
  p.w = d->w; p.bias = d->bias; p.N = d->N; p.K = d->ksize * d->ksize * d->C0 + p.C1;
  p.M = d->B * p.Ho * p.Wo;
  p.epi = d->epilogue; p.gamma = d->gamma; p.res = d->res; p.out = d->out; p.r = d->shuffle_r; p.a_act = d->a_act;
  if (p.M == 0) return 0;
  dim3 block(NT);
  if (p.N > 64) {
    dim3 grid((p.M + BM - 1) / BM, (p.N + 127) / 128);
    gemm_f32_kernel<128><<<grid, block, 0, stream>>>(p);
  } else if (p.N > 32) {
    dim3 grid((p.M + BM - 1) / BM, 1);
    gemm_f32_kernel<64><<<grid, block, 0, stream>>>(p);
  } else {
    dim3 grid((p.M + BM - 1) / BM, 1);
    gemm_f32_kernel<32><<<grid, block, 0, stream>>>(p);
  }
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

}  // namespace lvae
