/*
 * lvae_b200.h -- C ABI of the B200-native hierarchical-VAE rate-distortion path.
 *
 * The reference (duanzhiihao/lossy-vae) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md F1, section 8(b)); the drop-in boundary its callers see is the Python `lvae` package.
 * This header is the C boundary *underneath* that package: every entry point below replaces a
 * group of ATen/cuDNN/cuBLAS/CompressAI call sites of the reference, cited per function as
 * (reference file:line, relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are raw CUDA device addresses, `stream` is a
 *     cudaStream_t passed as void*; the caller owns every buffer; kernels never allocate.
 *   - activations are NHWC fp32: a feature map [B,H,W,C] is the row-major matrix [M=B*H*W, C].
 *   - every function returns 0 on success, a positive cudaError_t, or a negative LVAE_E_* code;
 *     lvae_last_error() returns a thread-local message.
 *   - no global mutable state besides one-time cudaFuncSetAttribute calls and the tensor-map
 *     driver entry point lookup.
 */
#ifndef LVAE_B200_H
#define LVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVAE_E_BADARG   (-1)
#define LVAE_E_UNSUPPORTED (-2)
#define LVAE_E_CORRUPT  (-3)

int lvae_version(void);
const char* lvae_last_error(void);

/* ---- GEMM-shaped operators -------------------------------------------------------------------
 * One descriptor covers every dense contraction of the path:
 *   nn.Linear fc1/fc2 of the ConvNeXt MLP          lvae/models/common.py:131-132,154
 *   patch_downsample (k = s = 2 | 4 conv)          common.py:29-30
 *   patch_upsample (1x1 conv + PixelShuffle)       common.py:33-38
 *   post_merge 1x1 on cat([feature, enc_feature])  lvae/models/qarv/model.py:36,66-67
 *   posterior 3x3 conv, prior / z_proj 1x1 convs   qarv/model.py:37-39,51,69,74
 * out[m, n] = epilogue( sum_k A[m, k] * Wp[n, k] + bias[n] )
 *   A[m, k]: m = (b, ho, wo) output pixel; k = (ky, kx, c) over segment 0 (NHWC a0 with C0 channels,
 *   square kernel `ksize`, `stride`, zero `pad`), followed for 1x1 ops by the C1 channels of
 *   segment 1 (a1) -- the K-concat that replaces torch.cat for post_merge.
 *   Wp is the packed weight [N, K] (K contiguous), produced by the host from the reference layout.
 */
enum lvae_epilogue {
  LVAE_EPI_BIAS = 0,          /* out = acc + bias                                   */
  LVAE_EPI_BIAS_GELU = 1,     /* out = gelu_erf(acc + bias)            (Mlp.act)     */
  LVAE_EPI_SCALE_RES = 2,     /* out = (acc + bias) * gamma[n] + res[m,n] (common.py:157-160) */
  LVAE_EPI_BIAS_RES = 3,      /* out = res[m,n] + (acc + bias)         (qarv/model.py:74) */
  LVAE_EPI_SHUFFLE_NHWC = 4,  /* PixelShuffle(r) into NHWC [B,H*r,W*r,N/r^2]; packed n = (i*r+j)*Co + c */
  LVAE_EPI_SHUFFLE_NCHW = 5,  /* PixelShuffle(r) into NCHW [B,N/r^2,H*r,W*r] (final image)  */
  LVAE_EPI_GELU_BWD = 6       /* out = (acc + bias) * gelu'(res[m,n]): data gradient through Mlp.act, res = the
                               * pre-activation (training step; tensor-core modes, plain [M,K] GEMMs, N % 4 == 0) */
};

enum lvae_precision {
  LVAE_PREC_FP32 = 0,         /* fp32 FFMA on CUDA cores                                                      */
  LVAE_PREC_BF16X3 = 1,       /* tcgen05 kind::f16, 2 bf16 planes per operand, 3 MMAs (hh + hm + mh): ~2^-17  */
  LVAE_PREC_BF16 = 2,         /* tcgen05, 1 plane, 1 MMA: ~2^-8 (non-parity fast mode)                        */
  LVAE_PREC_BF16X6 = 3,       /* tcgen05, 3 bf16 planes, 6 MMAs (hh + hm + mh + mm + hl + lh): ~2^-23, the    */
                              /* fp32-class mode in which quantised symbols match the fp32 CPU reference      */
  LVAE_PREC_F16 = 5,          /* tcgen05 kind::f16, ONE fp16 plane (weights scaled like F16X3), 1 MMA: ~2^-11. */
                              /* Not a parity mode for the coded part: the engine uses it ONLY for the blocks */
                              /* after CompresionStopFlag (qarv/zoo.py:78-88) in the 'f16x3+tail1' mode -- they */
                              /* cannot change a symbol or the rate, only the reconstruction (PSNR budget).   */
  LVAE_PREC_F16X3 = 4         /* tcgen05 kind::f16 on fp16 planes: 2 planes of 11 significand bits each, 3    */
                              /* MMAs (hh + hl + lh): ~2^-22 per product -- fp32-class at half the MMAs of    */
                              /* BF16X6.  Weight planes carry w * LVAE_F16_WEIGHT_SCALE (keeps the low plane  */
                              /* of small weights out of the fp16 subnormal range); the epilogue undoes it.   */
};
/* planes per operand for a precision mode: fp32 0, bf16 1, bf16x3 2, bf16x6 3, f16x3 2, f16 1 */
#define LVAE_MAX_PLANES 3
/* element format of operand planes */
enum lvae_plane_format { LVAE_PLANES_BF16 = 0, LVAE_PLANES_F16 = 1 };
#define LVAE_F16_WEIGHT_SCALE 256.0f

typedef struct lvae_gemm_desc {
  const float* a0;     /* segment 0 activations, NHWC [B,H,W,C0] */
  const float* a1;     /* segment 1 activations, NHWC [B,H,W,C1] or NULL */
  int32_t B, H, W;     /* input spatial dims */
  int32_t C0, C1;
  int32_t ksize, stride, pad;
  const float* w;      /* packed [N, K], K = ksize*ksize*C0 + C1 */
  const float* bias;   /* [N] or NULL */
  int32_t N;
  int32_t epilogue;    /* enum lvae_epilogue */
  const float* gamma;  /* [N]   (SCALE_RES) */
  const float* res;    /* [M,N] (SCALE_RES, BIAS_RES) */
  float* out;
  int32_t shuffle_r;   /* r for the SHUFFLE epilogues */
  int32_t precision;   /* enum lvae_precision */
  /* operand planes for the tensor-core path (device pointers; unused in fp32 mode).  A value x travels as the
   * bf16 planes p0 = rn(x), p1 = rn(x - p0), p2 = rn(x - p0 - p1) (lvae_split_bf16); a mode reads its first
   * 1 / 2 / 3 planes. */
  const void* w_planes[3];    /* bf16 [N,K] planes of w */
  const void* a_planes[3];    /* optional pre-split A planes (written by lvae_dwconv_ln_adaln_planes or by a previous
                               * lvae_gemm through out_planes); when [0] is set, a0/a1 are not read.  ksize == 1: plain
                               * [M,K] planes.  ksize == 3, stride 1, pad 1, C0 % 16 == 0, LVAE_EPI_BIAS or _BIAS_GELU:
                               * NHWC [B,H,W,C0] planes, convolved implicitly (9 shifted TMA box loads per channel
                               * block, no im2col workspace); the result may go to `out`, `out_planes` or both */
  void* out_planes[3];        /* optional bf16 [M,N] planes of the epilogue result (non-shuffle epilogues); `out`
                               * may then be NULL */
  void* workspace;            /* device scratch >= lvae_gemm_workspace_bytes(): im2col + split of a0/a1 when
                               * a_planes[0] == NULL */
  int64_t workspace_bytes;
  int32_t a_act;              /* 0: none; 1: GELU (erf form) applied to every element of a0 / a1 as it is read --
                               * VDBlock's c_i(gelu(x)) (lvae/models/qresvae/model.py:143-149).  Requires a0
                               * (not a_planes). */
  int32_t out_planes_act;     /* 1: out_planes hold gelu(result) instead of the result (the consumer is a VDBlock, whose first
                               * conv reads gelu(x), qresvae/model.py:144); `out` still receives the plain result */
  const void* a1_planes[3];   /* optional pre-split planes [M, C1] of segment 1 (the K-concat of post_merge): with
                               * a_planes (then [M, C0]) the tensor-core path reads both segments without an im2col /
                               * split pass.  ksize == 1, C0 % 64 == 0. */
} lvae_gemm_desc;

int lvae_gemm(const lvae_gemm_desc* d, void* stream);

/* The posterior head and the latent arithmetic of one layer in ONE launch (eval / compress branch of VRLVBlockBase.forward,
 * lvae/models/qarv/model.py:66-74 + 95-96 + 106-108): `d` describes the implicit 3x3 posterior convolution (pre-split A planes,
 * stride 1, pad 1, LVAE_EPI_BIAS, N = zdim <= 128, fp32 output); its result qm never reaches memory -- the epilogue reads the
 * prior parameters (pm | plogv_raw) [M, 2 zdim], and writes z = rint(qm - pm) + pm to d->out, the per-image sums of -ln P to
 * kl_partial[image * kl_stride + slot] (lvae_gemm_latent_num_partials(H, W, N) slots per image, one per (pixel tile, 32-column
 * chunk, row quarter): a fixed layout, so the sums are deterministic and batch-invariant; slots a launch does not own are left
 * untouched), optionally -ln P per element, and for the coder int32 symbols and scale-table indexes in NCHW order.  Same
 * arithmetic, bit for bit, as lvae_latent_eval (csrc/latent_math.cuh).  cdf_kind: LVAE_CDF_*. */
typedef struct lvae_latent_epilogue {
  const float* prior;        /* [M, 2 N]: pm | plogv_raw */
  const float* scale_table;  /* [n_scales] (needed with sym / idx) */
  int32_t n_scales, cdf_kind;
  float* kl_partial; int64_t kl_stride;
  float* kl_elem;            /* [M, N] or NULL */
  int32_t* sym; int32_t* idx;   /* [B, N, H, W] or both NULL */
} lvae_latent_epilogue;
int lvae_gemm_latent(const lvae_gemm_desc* d, const lvae_latent_epilogue* e, void* stream);
int lvae_gemm_latent_num_partials(int H, int W, int N);
int64_t lvae_gemm_workspace_bytes(const lvae_gemm_desc* d);
/* diagnostics: number of lvae_gemm calls served so far by the CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x BN
 * tiles; csrc/gemm2_tc.cu).  Opt-in (environment, read per call: LVAE_GEMM2=1; LVAE_G2_MIN_TILES / LVAE_G2_BN override
 * the eligibility threshold / the tile width): lvae_gemm then uses it for the large plain [M,K] x [N,K]^T contractions of
 * the 2-plane modes.  Results are bit-identical to the 128-row kernel; on B200 it measured slower, so it is off by
 * default (profiles/r2_gemm2_pair_kernel.md). */
long long lvae_gemm2_launch_count(void);
/* diagnostics: cycle breakdown (clock64) of CTA 0 of the LAST tensor-core GEMM launch -- which = 1: gemm_tc_kernel,
 * 2: the CTA-pair kernel.  out16[0] MMA-issuing thread total, [1] of it waiting for operand stages (TMA / L2), [2]
 * waiting for the epilogue to free the accumulator, [3] producer total, [4] producer waiting for a free stage,
 * [5..7] (pair kernel) one epilogue warp: total, waiting for the accumulator, draining TMEM; [8] tiles.  Synchronises. */
int lvae_debug_prof(int which, unsigned long long* out16);
/* Tuning knobs of the tensor-core GEMM (diagnostics / A-B measurements; results never depend on them): which 0 = force the
 * N-tile width BN (multiple of 16, 0 = automatic; env LVAE_TC_BN), 1 = unused,
 * 2 = narrower N-tiles for GEMMs that do not fill the SMs (default 1; LVAE_TC_AUTOBN). */
int lvae_set_tuning(int which, int value);
/* Host-only query (no device needed): the N-tile width BN a plain [M, K] x [N, K] tensor-core GEMM of this precision mode is
 * launched with on a device of n_sm SMs -- the widest tile that divides N evenly, or, for a 2-plane GEMM that would fill at most
 * two waves, the width that minimises waves x (K / 16) x (128 + 1.25 BN) cycles.  Results never depend on it. */
int lvae_gemm_tile_width(int M, int N, int K, int precision, int n_sm);
/* split fp32 -> bf16 planes p0 = rn(x), p1 = rn(x - p0), p2 = rn(x - p0 - p1); p1 / p2 may be NULL */
int lvae_split_bf16(const float* x, void* p0, void* p1, void* p2, int64_t n, void* stream);
/* the same split of (x * scale) into planes of `plane_format` (enum lvae_plane_format); weights of an
 * LVAE_PREC_F16X3 GEMM are split with plane_format = LVAE_PLANES_F16, scale = LVAE_F16_WEIGHT_SCALE, activations with
 * scale = 1.  fp16 conversion saturates at +-65504.  n % 2 == 0. */
int lvae_split_planes(const float* x, void* p0, void* p1, void* p2, int64_t n, int plane_format, float scale,
                      void* stream);
/* planes of gelu(x) (erf form): the operand of a VDBlock's first conv, computed once per encoder feature */
int lvae_gelu_split_planes(const float* x, void* p0, void* p1, void* p2, int64_t n, int plane_format, void* stream);

/* Fused ConvNeXt MLP (common.py:154-160) for narrow layers: out = res + gamma * (W2 gelu(W1 a + b1) + b2) in one
 * kernel; the hidden activation stays on chip.  a_p*: the two 16-bit planes [M, C] of the MLP input (written by
 * lvae_dwconv_ln_adaln_planes), w1_p* / w2_p*: planes of fc1.weight [hidden, C] / fc2.weight [C, hidden]
 * (lvae_split_planes; fp16 planes carry LVAE_F16_WEIGHT_SCALE).  C in {64, 128, 192}, hidden % 32 == 0, precision
 * LVAE_PREC_F16X3 or LVAE_PREC_BF16X3.  Bit-identical to the two lvae_gemm calls it replaces.  out may alias res. */
int lvae_convnext_mlp(const void* a_p0, const void* a_p1, const void* w1_p0, const void* w1_p1, const float* b1,
                      const void* w2_p0, const void* w2_p1, const float* b2, const float* gamma,
                      const float* res, float* out, int64_t M, int C, int hidden, int precision, void* stream);
/* the same, additionally writing the two 16-bit planes [M, C] of the result (out_planes_gelu = 0) or of gelu(result)
 * (out_planes_gelu = 1) for a tensor-core consumer */
int lvae_convnext_mlp_planes(const void* a_p0, const void* a_p1, const void* w1_p0, const void* w1_p1, const float* b1,
                             const void* w2_p0, const void* w2_p1, const float* b2, const float* gamma,
                             const float* res, float* out, void* out_p0, void* out_p1, int out_planes_gelu,
                             int64_t M, int C, int hidden, int precision, void* stream);

/* ---- depthwise conv + LayerNorm + AdaLN (common.py:145-152) ------------------------------------
 * y[m, c] = LN_c( dwconv_kxk(x)[m, c] + dw_bias[c] ) * (1 + scale[b, c]) + shift[b, c]
 * x NHWC [B,H,W,C]; dw_w packed [k*k, C]; ada = [B, ada_stride] with shift at ada[b, ada_off + c]
 * and scale at ada[b, ada_off + C + c]; y row-major [M, C]. C % 64 == 0, k in {1,3,5,7}.
 * If ln_w != NULL the affine LayerNorm of qresvae's block is applied instead of AdaLN
 * (lvae/models/qresvae/model.py:163-182). */
int lvae_dwconv_ln_adaln(const float* x, const float* dw_w, const float* dw_b,
                         const float* ada, int64_t ada_stride, int64_t ada_off,
                         const float* ln_w, const float* ln_b,
                         float* y, int B, int H, int W, int C, int k, void* stream);
/* Same operator writing the result as 16-bit planes [M, C] (enum lvae_plane_format) -- the A operand of the
 * tensor-core fc1 GEMM (y1 / y2 may be NULL when the precision mode reads fewer planes). */
int lvae_dwconv_ln_adaln_planes(const float* x, const float* dw_w, const float* dw_b,
                                const float* ada, int64_t ada_stride, int64_t ada_off,
                                const float* ln_w, const float* ln_b,
                                void* y0, void* y1, void* y2, int plane_format,
                                int B, int H, int W, int C, int k, void* stream);

/* ---- backward of the dwconv + LayerNorm + modulation stage (training step; autograd of common.py:145-152 and of
 * qresvae/model.py:163-182, which the reference gets from ATen) -------------------------------------------------
 * lvae_dwconv: y = dwconv_kxk(x) [+ bias[c]] [+ add], NHWC fp32, filter packed [k*k, C] (bias / add may be NULL).
 * flip = 0: the forward convolution (recomputes the conv output c for the LayerNorm backward); flip = 1: correlation with
 * the spatially flipped filter = the data gradient dx = conv^T(dc), `add` carrying the residual branch's gradient. */
int lvae_dwconv(const float* x, const float* dw_w, const float* bias, const float* add, float* y,
                int B, int H, int W, int C, int k, int flip, void* stream);
/* dw[t, c] = sum_p dc[p, c] * x[p + t, c] (packed [k*k, C]) and db[c] = sum_p dc[p, c]; both outputs are zeroed first,
 * partial sums meet in fp32 atomics (order not fixed, like cuDNN's weight gradients). */
int lvae_dwconv_wgrad(const float* dc, const float* x, float* dw, float* db,
                      int B, int H, int W, int C, int k, void* stream);
/* LayerNorm(C, eps 1e-6) + modulation a = yhat * g1 + g0 backward: from the conv output c [M,C] and da = dL/da [M,C],
 * dc [M,C] = dL/dc and dmod = (sum da | sum da * yhat): one [2C] row per image for AdaLN (g1 = 1 + scale[b], g0 = shift[b]
 * read from ada as in lvae_dwconv_ln_adaln: dmod row = (dshift | dscale)), a single row (d ln_b | d ln_w) when ln_w != NULL. */
int lvae_ln_mod_bwd(const float* c, const float* da, const float* ada, int64_t ada_stride, int64_t ada_off,
                    const float* ln_w, float* dc, float* dmod, int B, int HW, int C, void* stream);

/* ---- weight gradients of the dense layers (training step; autograd of F.linear / 1x1 conv2d) -------------------
 * dW[n, k] = sum_p dY[p, n] * X[p, k] on the tensor cores: lvae_split_planes_t transposes a pixel-major fp32 matrix
 * [P, C] into two K-major bf16 planes [C, P] (hi | lo; P even), lvae_gemm_wgrad contracts two such operands over P
 * (P % 8 == 0) with split-K over the grid and writes dw [n_out, k_in] fp32 (zeroed first, fp32 atomics). */
int lvae_split_planes_t(const float* x, void* p0, void* p1, int64_t P, int C, void* stream);
/* the same with act = 1: planes of gelu(x) (the fc2 weight gradient's operand from the recomputed pre-activation), and /
 * or colsum != NULL: colsum[c] += sum_p x[p, c] (the bias gradient of the layer whose dY is being split; the caller
 * zeroes it) */
int lvae_split_planes_t_ex(const float* x, void* p0, void* p1, int64_t P, int C, int act, float* colsum, void* stream);
/* The same K-major bf16 planes [C, P] from an operand that already exists as two 16-bit planes [P, C] (plane_format: LVAE_PLANES_*):
 * value = hi + lo, re-split.  Lets the fc1 weight gradient reuse the recomputed forward operand instead of an fp32 recomputation. */
int lvae_planes_transpose(const void* a0, const void* a1, int plane_format, void* p0, void* p1, int64_t P, int C, void* stream);
int lvae_gemm_wgrad(const void* dyt_p0, const void* dyt_p1, const void* xt_p0, const void* xt_p1,
                    float* dw, int n_out, int k_in, int64_t P, void* stream);

/* ---- training-step tail on flat buffers: global-norm clipping + Adam + EMA (lvae/trainer.py:360-377,394-406) --------
 * Replaces clip_grad_norm_(params, max_norm) -> torch.optim.Adam.step() (weight_decay 0, no amsgrad) -> ModelEmaV2.update
 * when parameters, gradients, Adam moments and the EMA copy each live in ONE flat fp32 buffer of n elements (16-byte
 * aligned; lvae.training.GraphedTrainStep lays them out).  Two launches, 36 bytes per parameter:
 *   g' = g * min(1, max_norm / (||g||_2 + 1e-6))   (max_norm <= 0: no clipping)
 *   m += (g' - m)(1 - beta1);  v = beta2 v + (1 - beta2) g'^2;  p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
 *   ema = decay * ema + (1 - decay) * p            (timm ModelEmaV2, same roundings; ema == NULL: skipped)
 * lr, step (= t of this update, >= 1, as a float) and ema_decay (TWO floats: decay and 1 - decay, rounded from the host's
 * double values) are DEVICE scalars, so that a captured step can be
 * replayed while the host runs the learning-rate schedule / EMA warm-up.  scratch: lvae_optim_scratch_doubles() doubles
 * (deterministic two-stage sum of squares).  grad_norm_out (optional, device): ||g||_2 before clipping. */
int lvae_optim_scratch_doubles(void);
int lvae_adam_clip_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, double* scratch,
                       float max_norm, const float* lr, const float* step, const float* ema_decay,
                       double beta1, double beta2, double eps, float* grad_norm_out, void* stream);

/* ---- GaussianNLLOutputNet of qres34m_lossless (qresvae/model.py:16-94): the arithmetic after the two patch_upsample
 * heads (which are lvae_gemm launches with LVAE_EPI_SHUFFLE_NCHW).  All tensors NCHW [B,3,H,W] fp32, chw = 3*H*W.
 * lvae_nll_output: nll_partial[b * np + blk] += -gaussian_log_prob_mass(p_mean, exp(softplus(p_logscale + 16) - 16),
 *   (im - .5) * 2, bin = 1/127.5, prob_clamp = 1e-6) summed per block, np = lvae_image_num_partials(chw) (model.py:24-40).
 * lvae_outnet_codec: pm = (round(p_mean*127.5+127.5)/127.5 - 1)/bin, idx = table index of max(exp(p_logscale - ln bin),
 *   0.11) on scale_table (n_scales <= 128) and, with im != NULL, sym = round((im - .5)*2/bin - pm) (model.py:68-86).
 * lvae_outnet_decode: im_hat = clamp((sym + pm) * bin, -1, 1) * .5 + .5 (model.py:88-94 + process_output). */
int lvae_nll_output(const float* p_mean, const float* p_logscale, const float* im, float* nll_partial,
                    int B, int chw, void* stream);
int lvae_outnet_codec(const float* p_mean, const float* p_logscale, const float* im, const float* scale_table,
                      int n_scales, float* pm, int32_t* idx, int32_t* sym, int64_t total, void* stream);
int lvae_outnet_decode(const int32_t* sym, const float* pm, float* im_hat, int64_t total, void* stream);

/* GPU-side input pipeline of the training step: RandomCrop(crop) + RandomHorizontalFlip + ToTensor
 * (lvae/datasets/image.py:45-56) on a device batch of decoded uint8 images src [B,3,Hs,Ws]; y0 / x0 [B] int32 crop
 * origins (y0 + crop <= Hs, x0 + crop <= Ws), flip [B] uint8 -> out [B,3,crop,crop] fp32 = x / 255. */
int lvae_crop_flip_u8(const void* src, const int32_t* y0, const int32_t* x0, const void* flip, float* out,
                      int B, int Hs, int Ws, int crop, void* stream);

/* ---- fused latent-layer kernels (qarv/model.py:51-53,90-96,104-113; CompressAI GaussianConditional) --
 * prior [M, 2*zdim] (pm | plogv_raw) and qm [M, zdim] are NHWC matrices; hw = h*w positions per image.
 * Eval (K11+K12+K15): z = rint(qm-pm)+pm; kl = -ln max(Phi((.5-|z-pm|)/s) - Phi((-.5-|z-pm|)/s), 1e-9),
 * s = max(exp(softplus(plogv+2.3)-2.3), 0.11).  kl_partial[b * kl_stride + blk] gets a deterministic
 * partial sum (blk < lvae_latent_num_partials() <= kl_stride), so the layers of a model can share one
 * [B, kl_stride] matrix, each at its own column offset.  Optional outputs (NULL to skip): kl_elem [M,zdim];
 * sym (int32) and idx (int32) in NCHW order [B,zdim,h,w] for the host coder (K14). */
int lvae_latent_num_partials(int hw, int zdim);
/* cdf_kind selects how the standard normal CDF is evaluated, because that is where the two reference families differ:
 * LVAE_CDF_NORMAL: 0.5 * (1 + erf(t / sqrt 2)) = td.Normal(0,1).cdf, the override of qarv's DiscretizedGaussian
 * (lvae/models/entropy_coding.py:77-82); LVAE_CDF_ERFC: 0.5 * erfc(-t / sqrt 2), CompressAI's
 * GaussianConditional._standardized_cumulative, which qres34m uses unmodified (lvae/models/qresvae/model.py:241). */
enum lvae_cdf_kind { LVAE_CDF_NORMAL = 0, LVAE_CDF_ERFC = 1 };
int lvae_latent_eval(const float* qm, const float* prior, const float* scale_table, int n_scales,
                     float* z, float* kl_partial, int kl_stride, float* kl_elem, int32_t* sym, int32_t* idx,
                     int B, int hw, int zdim, int cdf_kind, void* stream);
/* Training (K13): z = qm + noise; kl = -gaussian_log_prob_mass(pm, pv, z) (entropy_coding.py:17-49) */
int lvae_latent_train(const float* qm, const float* prior, const float* noise,
                      float* z, float* kl_partial, int kl_stride, float* kl_elem,
                      int B, int hw, int zdim, void* stream);
/* Backward of lvae_latent_train (entropy_coding.py:17-49 under autograd).  Upstream gradients:
 * dz [M,zdim] = dL/dz arriving through z_proj (NULL: 0) and dL/dkl, either per element (dkl_elem [M,zdim]) or the
 * constant dkl_scale (the loss uses 1 / (ndims * B), qarv/model.py:338-346).  Outputs: dqm [M,zdim] = dL/dqm and
 * dprior [M,2*zdim] = (dL/dpm | dL/dplogv_raw), the gradient w.r.t. the prior head's output. */
int lvae_latent_train_bwd(const float* qm, const float* prior, const float* noise,
                          const float* dz, const float* dkl_elem, float dkl_scale,
                          float* dqm, float* dprior, int B, int hw, int zdim, void* stream);
/* Decompress side: idx from the prior only (NCHW int32), then z = float(sym) + pm from decoded symbols */
int lvae_latent_prior_index(const float* prior, const float* scale_table, int n_scales,
                            int32_t* idx, int B, int hw, int zdim, void* stream);
int lvae_latent_dequant(const int32_t* sym, const float* prior, float* z,
                        int B, int hw, int zdim, void* stream);
/* Sampling (qarv/model.py:98-103): z = pm + pv*randn*t + uniform*t with caller-supplied noise */
int lvae_latent_sample(const float* prior, const float* randn, const float* unif, float t,
                       float* z, int B, int hw, int zdim, void* stream);

/* rd model (lvae/models/rd/model.py:27-49,162-227): continuous Gaussian posterior.  post / prior are [M, 2*zdim]
 * (mean_raw | std_raw); mean = linear_sqrt(raw), std = softplus(raw, beta = ln 2, threshold = 12);
 * kl = gaussian_kl(qm, qv, pm, pv); z = qm + qv * noise with caller-supplied N(0,1) noise [M, zdim].
 * kl_partial as for lvae_latent_eval. */
int lvae_rd_latent(const float* post, const float* prior, const float* noise,
                   float* z, float* kl_partial, int kl_stride, float* kl_elem,
                   int B, int hw, int zdim, void* stream);
/* z = pm + pv * randn * t (rd/model.py:216) */
int lvae_rd_sample(const float* prior, const float* randn, float t, float* z,
                   int B, int hw, int zdim, void* stream);

/* ---- small host-side-M operators ----------------------------------------------------------------
 * lmb -> sinusoidal embedding (common.py:101-107, qarv/model.py:275-287): emb0[b, :] =
 * [cos(a*f) | sin(a*f)], a = log(lmb[b]) * period / log(max_lmb), f = host-supplied [dim/2] table. */
int lvae_lmb_sinusoid(const float* lmb, const float* freqs, float* emb0, int B, int dim,
                      float period, float max_lmb, void* stream);
/* out[b, n] = act_out( sum_k act_in(x[b,k]) * w[n,k] + bias[n] ); act flags: 0 none, 1 gelu_erf.
 * Used for the lambda-embedding MLP and for ALL AdaLN projections of a model in one launch
 * (common.py:150: embedding_layer = GELU -> Linear(256, 2C)). */
int lvae_small_linear(const float* x, const float* w, const float* bias, float* out,
                      int B, int K, int N, int act_in, int act_out, void* stream);

/* ---- image side (qarv/model.py:213-242,338-346) ---------------------------------------------------
 * im NCHW [B,3,H,W] in [0,1] -> space-to-depth operand [B*H/4*W/4, 48], k = (i*4+j)*3 + c,
 * value (im + shift) * scale */
int lvae_image_to_patches(const float* im, float* a, int B, int H, int W, int r,
                          float shift, float scale, void* stream);
/* per-image partial sums: sq_target[b] = sum (x_hat - (im-.5)*2)^2, sq_im[b] = sum (clamp(x_hat)*.5+.5 - im)^2;
 * also writes im_hat if non-NULL. partial arrays are [B, lvae_image_num_partials()] */
int lvae_image_num_partials(int chw);
int lvae_image_distortion(const float* x_hat, const float* im, float* im_hat,
                          float* sq_target_partial, float* sq_im_partial, int B, int chw, void* stream);
/* loss assembly (qarv/model.py:338-358): stats = [loss, mean_b kl/ndims (nats), mean_b mse, image-domain mse,
 * kl_img[B] (nats per dimension), mse_img[B], sqerr_img[B]]; kl_partial is [B, kl_stride] with kl_cols
 * valid columns; ndims = C*H*W of the image; lmb [B]. stats must hold 4 + 3*B floats. */
int lvae_rd_finalize(const float* kl_partial, int kl_stride, int kl_cols,
                     const float* sq_target_partial, const float* sq_im_partial, int np,
                     const float* lmb, int B, int64_t ndims, float* stats, void* stream);
/* feature[b,h,w,c] = bias[c] (qarv/model.py:289-292) */
int lvae_broadcast_bias(const float* bias, float* out, int64_t M, int C, void* stream);
/* dst[m, 0:C] = src[m, 0:C], dst[m, C:Cp] = 0: latent tensors whose channel count is not a multiple of 8 (qres34m
 * zdim 14 / 12 / 10) are padded before the 3x3 z_proj conv (lvae/models/qresvae/model.py:236-240) */
int lvae_pad_channels(const float* src, float* dst, int64_t M, int C, int Cp, void* stream);
/* out[i] = sum_j partial[i, j] in fixed order (double accumulation), n rows of `cols` */
int lvae_sum_partials(const float* partial, float* out, int n, int cols, void* stream);

/* ---- host entropy coder (CompressAI: entropy_models.py update()/compress()/decompress(),
 *      cpp_exts/ops/ops.cpp pmf_to_quantized_cdf, cpp_exts/rans/rans_interface.cpp; call sites
 *      qarv/model.py:106-113,123-124) -- host pointers only --------------------------------------- */
/* pmf (fp32, length n) -> quantized cdf (length n+1, last = 1<<precision) */
int lvae_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int32_t* cdf_out);
/* worst-case encoded size in bytes for n symbols */
int64_t lvae_rans_bound(int64_t n);
/* encode n symbols; cdf is [n_cdf, cdf_stride] int32; returns bytes written in *out_len */
int lvae_rans_encode(const int32_t* sym, const int32_t* idx, int64_t n,
                     const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, const int32_t* offset,
                     int n_cdf, uint8_t* out, int64_t out_cap, int64_t* out_len);
/* n_streams independent streams on up to n_threads host threads (largest streams should come first): stream i covers
 * symbols [begin[i], begin[i+1]) and writes out_len[i] bytes at out + out_begin[i] (capacity out_begin[i+1] -
 * out_begin[i] >= lvae_rans_bound of its symbol count).  SURVEY 8(f)-1. */
int lvae_rans_encode_streams(const int32_t* sym, const int32_t* idx, const int64_t* begin, int n_streams,
                             const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, const int32_t* offset,
                             int n_cdf, uint8_t* out, const int64_t* out_begin, int64_t* out_len, int n_threads);
/* the decoding counterpart: stream i = bytes [in_begin[i], in_begin[i+1]) of `in` -> symbols [begin[i], begin[i+1]) */
int lvae_rans_decode_streams(const uint8_t* in, const int64_t* in_begin, const int32_t* idx, const int64_t* begin,
                             int n_streams, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                             const int32_t* offset, int n_cdf, int32_t* sym_out, int n_threads);
int lvae_rans_decode(const uint8_t* in, int64_t in_len, const int32_t* idx, int64_t n,
                     const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, const int32_t* offset,
                     int n_cdf, int32_t* sym_out);

#ifdef __cplusplus
}
#endif
#endif /* LVAE_B200_H */
