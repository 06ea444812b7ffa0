"""ctypes binding of liblvae_b200.so (the C ABI declared in include/lvae_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent.parent / 'lib' / 'liblvae_b200.so'
_lib = None

EPI_BIAS, EPI_BIAS_GELU, EPI_SCALE_RES, EPI_BIAS_RES, EPI_SHUFFLE_NHWC, EPI_SHUFFLE_NCHW, EPI_GELU_BWD = range(7)
PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_BF16X6, PREC_F16X3, PREC_F16 = range(6)
PRECISIONS = {'fp32': PREC_FP32, 'bf16x3': PREC_BF16X3, 'bf16': PREC_BF16, 'bf16x6': PREC_BF16X6, 'f16x3': PREC_F16X3}
NUM_PLANES = {PREC_FP32: 0, PREC_BF16: 1, PREC_BF16X3: 2, PREC_BF16X6: 3, PREC_F16X3: 2, PREC_F16: 1}
MMA_TERMS = {PREC_FP32: 1, PREC_BF16: 1, PREC_BF16X3: 3, PREC_BF16X6: 6, PREC_F16X3: 3, PREC_F16: 1}
PLANES_BF16, PLANES_F16 = 0, 1
CDF_NORMAL, CDF_ERFC = 0, 1
PLANE_FORMAT = {PREC_FP32: PLANES_BF16, PREC_BF16: PLANES_BF16, PREC_BF16X3: PLANES_BF16, PREC_BF16X6: PLANES_BF16,
                PREC_F16X3: PLANES_F16, PREC_F16: PLANES_F16}
F16_WEIGHT_SCALE = 256.0      # LVAE_F16_WEIGHT_SCALE

_fp = C.c_void_p   # device / host pointers are passed as integers


class GemmDesc(C.Structure):
    _fields_ = [
        ('a0', _fp), ('a1', _fp),
        ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
        ('C0', C.c_int32), ('C1', C.c_int32),
        ('ksize', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32),
        ('w', _fp), ('bias', _fp),
        ('N', C.c_int32), ('epilogue', C.c_int32),
        ('gamma', _fp), ('res', _fp), ('out', _fp),
        ('shuffle_r', C.c_int32), ('precision', C.c_int32),
        ('w_planes', _fp * 3), ('a_planes', _fp * 3), ('out_planes', _fp * 3),
        ('workspace', _fp), ('workspace_bytes', C.c_int64),
        ('a_act', C.c_int32), ('out_planes_act', C.c_int32),
        ('a1_planes', _fp * 3),
    ]


class LatentEpilogue(C.Structure):
    """lvae_latent_epilogue (include/lvae_b200.h): the latent arithmetic as the epilogue of the posterior convolution"""
    _fields_ = [
        ('prior', _fp), ('scale_table', _fp),
        ('n_scales', C.c_int32), ('cdf_kind', C.c_int32),
        ('kl_partial', _fp), ('kl_stride', C.c_int64),
        ('kl_elem', _fp),
        ('sym', _fp), ('idx', _fp),
    ]


_PROTOS = {
    'lvae_version': (C.c_int, []),
    'lvae_last_error': (C.c_char_p, []),
    'lvae_gemm': (C.c_int, [C.POINTER(GemmDesc), _fp]),
    'lvae_gemm_workspace_bytes': (C.c_int64, [C.POINTER(GemmDesc)]),
    'lvae_gemm_latent': (C.c_int, [C.POINTER(GemmDesc), C.POINTER(LatentEpilogue), _fp]),
    'lvae_gemm_latent_num_partials': (C.c_int, [C.c_int, C.c_int, C.c_int]),
    'lvae_gemm2_launch_count': (C.c_longlong, []),
    'lvae_debug_prof': (C.c_int, [C.c_int, _fp]),
    'lvae_set_tuning': (C.c_int, [C.c_int, C.c_int]),
    'lvae_gemm_tile_width': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'lvae_convnext_mlp': (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_convnext_mlp_planes': (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int,
                                           C.c_int64, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_gelu_split_planes': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, _fp]),
    'lvae_split_bf16': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _fp]),
    'lvae_dwconv_ln_adaln': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int64, _fp, _fp, _fp,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_split_planes': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_float, _fp]),
    'lvae_dwconv_ln_adaln_planes': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int64, _fp, _fp, _fp, _fp, _fp, C.c_int,
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_dwconv': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_dwconv_wgrad': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_ln_mod_bwd': (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int64, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_split_planes_t': (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int, _fp]),
    'lvae_split_planes_t_ex': (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, _fp, _fp]),
    'lvae_planes_transpose': (C.c_int, [_fp, _fp, C.c_int, _fp, _fp, C.c_int64, C.c_int, _fp]),
    'lvae_gemm_wgrad': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int64, _fp]),
    'lvae_nll_output': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp]),
    'lvae_outnet_codec': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, _fp, _fp, _fp, C.c_int64, _fp]),
    'lvae_outnet_decode': (C.c_int, [_fp, _fp, _fp, C.c_int64, _fp]),
    'lvae_crop_flip_u8': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_optim_scratch_doubles': (C.c_int, []),
    'lvae_adam_clip_ema': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, _fp, C.c_float, _fp, _fp, _fp,
                                     C.c_double, C.c_double, C.c_double, _fp, _fp]),
    'lvae_latent_num_partials': (C.c_int, [C.c_int, C.c_int]),
    'lvae_latent_eval': (C.c_int, [_fp, _fp, _fp, C.c_int, _fp, _fp, C.c_int, _fp, _fp, _fp,
                                   C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_pad_channels': (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int, _fp]),
    'lvae_latent_train': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_latent_train_bwd': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_float, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_latent_prior_index': (C.c_int, [_fp, _fp, C.c_int, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_latent_dequant': (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_latent_sample': (C.c_int, [_fp, _fp, _fp, C.c_float, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_rd_latent': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_rd_sample': (C.c_int, [_fp, _fp, C.c_float, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_lmb_sinusoid': (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, _fp]),
    'lvae_small_linear': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    'lvae_image_to_patches': (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp]),
    'lvae_image_num_partials': (C.c_int, [C.c_int]),
    'lvae_image_distortion': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp]),
    'lvae_rd_finalize': (C.c_int, [_fp, C.c_int, C.c_int, _fp, _fp, C.c_int, _fp, C.c_int, C.c_int64, _fp, _fp]),
    'lvae_broadcast_bias': (C.c_int, [_fp, _fp, C.c_int64, C.c_int, _fp]),
    'lvae_sum_partials': (C.c_int, [_fp, _fp, C.c_int, C.c_int, _fp]),
    'lvae_pmf_to_quantized_cdf': (C.c_int, [_fp, C.c_int, C.c_int, _fp]),
    'lvae_rans_bound': (C.c_int64, [C.c_int64]),
    'lvae_rans_encode': (C.c_int, [_fp, _fp, C.c_int64, _fp, C.c_int, _fp, _fp, C.c_int, _fp, C.c_int64,
                                   C.POINTER(C.c_int64)]),
    'lvae_rans_encode_streams': (C.c_int, [_fp, _fp, _fp, C.c_int, _fp, C.c_int, _fp, _fp, C.c_int, _fp, _fp, _fp, C.c_int]),
    'lvae_rans_decode_streams': (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, _fp, C.c_int, _fp, _fp, C.c_int, _fp, C.c_int]),
    'lvae_rans_decode': (C.c_int, [_fp, C.c_int64, _fp, C.c_int64, _fp, C.c_int, _fp, _fp, C.c_int, _fp]),
}

EXPORTS = tuple(_PROTOS)


def lib_path():
    return _LIB_PATH


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.is_file():
            raise RuntimeError(
                f'{_LIB_PATH} not found: the sm_100a CUDA library has not been built. '
                f'Run `python lossy-vae_b200/build.py` (or __graft_entry__.build()). There is no CPU fallback.')
        L = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what=''):
    if code != 0:
        msg = lib().lvae_last_error().decode(errors='replace')
        raise RuntimeError(f'liblvae_b200 {what} failed with code {code}: {msg}')


def set_planes(desc, which, tensors):
    """desc.<which>_planes[i] = tensors[i].data_ptr() (which in 'w', 'a', 'out'); missing planes are NULL."""
    arr = getattr(desc, which + '_planes')
    for i in range(3):
        arr[i] = tensors[i].data_ptr() if tensors is not None and i < len(tensors) and tensors[i] is not None else None


launch_count = 0   # number of kernel-launching C calls issued (bench.py reports it)
