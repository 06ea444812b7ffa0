// extern "C" entry for the GEMM-shaped operators: argument validation and dispatch between the
// fp32 CUDA-core kernel (gemm_f32.cu) and the tcgen05 tensor-core kernel (gemm_tc.cu).
#include "common.cuh"

namespace lvae {
int gemm_f32_launch(const lvae_gemm_desc* d, cudaStream_t stream);
int gemm_tc_launch(const lvae_gemm_desc* d, cudaStream_t stream);
int gemm_tc_launch_latent(const lvae_gemm_desc* d, const lvae_latent_epilogue* lat, cudaStream_t stream);
int gemm_tc_latent_num_partials(int H, int W, int N);
int gemm_tc_tile_width(int M, int N, int K, int precision, int n_sm);
bool gemm_smallk_applicable(const lvae_gemm_desc* d);
int gemm_smallk_launch(const lvae_gemm_desc* d, cudaStream_t stream);
int64_t gemm_tc_workspace_bytes(const lvae_gemm_desc* d);
}

extern "C" int64_t lvae_gemm_workspace_bytes(const lvae_gemm_desc* d) {
  if (!d || d->precision == LVAE_PREC_FP32) return 0;
  return lvae::gemm_tc_workspace_bytes(d);
}

extern "C" int lvae_gemm_tile_width(int M, int N, int K, int precision, int n_sm) {
  if (M <= 0 || N <= 0 || K <= 0 || n_sm <= 0 || precision == LVAE_PREC_FP32) return LVAE_E_BADARG;
  return lvae::gemm_tc_tile_width(M, N, K, precision, n_sm);
}

extern "C" int lvae_gemm_latent_num_partials(int H, int W, int N) { return lvae::gemm_tc_latent_num_partials(H, W, N); }

extern "C" int lvae_gemm_latent(const lvae_gemm_desc* d, const lvae_latent_epilogue* e, void* stream) {
  using namespace lvae;
  LVAE_CHECK_ARG(d != nullptr && e != nullptr && d->w != nullptr && d->a_planes[0] != nullptr && d->out != nullptr);
  LVAE_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->N > 0 && d->C0 > 0);
  LVAE_CHECK_ARG(d->ksize == 3 && d->stride == 1 && d->pad == 1 && d->a1 == nullptr && d->a1_planes[0] == nullptr);
  LVAE_CHECK_ARG(d->precision != LVAE_PREC_FP32);
  return gemm_tc_launch_latent(d, e, (cudaStream_t)stream);
}

extern "C" int lvae_gemm(const lvae_gemm_desc* d, void* stream) {
  using namespace lvae;
  LVAE_CHECK_ARG(d != nullptr);
  LVAE_CHECK_ARG(d->w != nullptr);
  LVAE_CHECK_ARG(d->a0 != nullptr || (d->a_planes[0] != nullptr && d->precision != LVAE_PREC_FP32));
  LVAE_CHECK_ARG(d->out != nullptr || (d->out_planes[0] != nullptr && d->precision != LVAE_PREC_FP32));
  LVAE_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->N > 0);
  LVAE_CHECK_ARG(d->C0 > 0 && d->C0 % 4 == 0);
  LVAE_CHECK_ARG(d->a1 == nullptr || d->a_planes[0] != nullptr || (d->C1 > 0 && d->C1 % 4 == 0 && d->ksize == 1 && d->stride == 1 && d->pad == 0));
  LVAE_CHECK_ARG(d->ksize >= 1 && d->stride >= 1 && d->pad >= 0);
  LVAE_CHECK_ARG(d->out_planes_act == 0 || (d->out_planes_act == 1 && d->out_planes[0] != nullptr && d->precision != LVAE_PREC_FP32 &&
                                            d->epilogue != LVAE_EPI_BIAS_GELU && d->N % 4 == 0 && d->ksize == 1));
  LVAE_CHECK_ARG(d->a_act == 0 || (d->a_act == 1 && d->a0 != nullptr && d->a_planes[0] == nullptr));
  LVAE_CHECK_ARG(d->H + 2 * d->pad >= d->ksize && d->W + 2 * d->pad >= d->ksize);
  LVAE_CHECK_ARG(d->epilogue >= LVAE_EPI_BIAS && d->epilogue <= LVAE_EPI_GELU_BWD);
  if (d->epilogue == LVAE_EPI_GELU_BWD)
    LVAE_CHECK_ARG(d->res != nullptr && d->out != nullptr && d->precision != LVAE_PREC_FP32 && d->ksize == 1 && d->N % 4 == 0 &&
                   d->C0 >= 64);        // out_planes: operand planes of the result as well (the next data-gradient GEMM's A)
  if (d->epilogue == LVAE_EPI_SCALE_RES) LVAE_CHECK_ARG(d->gamma && d->res);
  if (d->epilogue == LVAE_EPI_BIAS_RES) LVAE_CHECK_ARG(d->res != nullptr);
  if (d->epilogue == LVAE_EPI_SHUFFLE_NHWC || d->epilogue == LVAE_EPI_SHUFFLE_NCHW)
    LVAE_CHECK_ARG(d->shuffle_r >= 1 && d->N % (d->shuffle_r * d->shuffle_r) == 0 && d->out != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  // rank-K updates with tiny K (z_proj) are pure bandwidth: exact fp32 FMAs on the CUDA cores in every mode
  if (gemm_smallk_applicable(d)) return gemm_smallk_launch(d, st);
  switch (d->precision) {
    case LVAE_PREC_FP32: return gemm_f32_launch(d, st);
    case LVAE_PREC_BF16X3:
    case LVAE_PREC_BF16:
    case LVAE_PREC_BF16X6:
    case LVAE_PREC_F16:
    case LVAE_PREC_F16X3: return gemm_tc_launch(d, st);
    default: set_error("unknown precision mode %d", d->precision); return LVAE_E_BADARG;
  }
}
