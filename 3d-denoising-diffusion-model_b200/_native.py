"""ctypes binding of libddpm3d.so (include/ddpm3d.h).  There is no fallback:
if the library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DDPM3D_LIB: another build of the same library (same-box A/B of two kernel versions; tools only)
LIB_PATH = os.environ.get("DDPM3D_LIB") or os.path.join(HERE, "libddpm3d.so")

FP32, BF16, FP16, BF16_STRICT = 0, 1, 2, 3
MEAN_PREVIOUS_X, MEAN_START_X, MEAN_EPSILON = 0, 1, 2
VAR_LEARNED, VAR_FIXED_SMALL, VAR_FIXED_LARGE, VAR_LEARNED_RANGE = 0, 1, 2, 3
MAX_LEVELS = 8


class Config(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("in_channels", C.c_int32), ("model_channels", C.c_int32),
        ("out_channels", C.c_int32), ("num_res_blocks", C.c_int32), ("n_levels", C.c_int32),
        ("channel_mult", C.c_int32 * MAX_LEVELS), ("n_attention_ds", C.c_int32),
        ("attention_ds", C.c_int32 * MAX_LEVELS), ("num_classes", C.c_int32), ("num_heads", C.c_int32),
        ("num_head_channels", C.c_int32), ("num_heads_upsample", C.c_int32),
        ("use_scale_shift_norm", C.c_int32), ("resblock_updown", C.c_int32),
        ("use_new_attention_order", C.c_int32), ("precision", C.c_int32),
        ("dims", C.c_int32), ("middle_attention", C.c_int32), ("unconditional", C.c_int32),
    ]


class StepScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "model_t", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
        "posterior_mean_coef2", "min_log", "max_log", "fixed_variance", "fixed_log_variance",
        "recip_coef1", "coef2_over_coef1", "alphas_cumprod", "alphas_cumprod_prev", "pad0_", "pad1_", "pad2_")]


class ProfRecord(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad_", C.c_int32), ("ms", C.c_float), ("pad2_", C.c_float),
                ("work", C.c_double)]


PROF_KINDS = ["conv_tcgen05", "conv_simt", "gn_stats", "gn_finalize", "groupnorm", "embedding", "update",
              "attention", "misc", "conv_small", "halo_exchange", "gn_allgather", "empty_bracket"]

_P = C.c_void_p
_I = C.c_int
# name -> (restype, argtypes); every symbol include/ddpm3d.h declares
SIGNATURES = {
    "ddpm3d_last_error": (C.c_char_p, []),
    "ddpm3d_abi_version": (_I, []),
    "ddpm3d_create": (_I, [C.POINTER(Config), C.POINTER(_P)]),
    "ddpm3d_destroy": (None, [_P]),
    "ddpm3d_param_count": (_I, [_P]),
    "ddpm3d_param_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(_I)]),
    "ddpm3d_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "ddpm3d_finalize_weights": (_I, [_P, _I]),
    "ddpm3d_set_timestep_freqs": (_I, [_P, _P, _I]),
    "ddpm3d_workspace_bytes": (C.c_int64, [_P, _I, _I, _I, _I]),
    "ddpm3d_unet_forward": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ddpm3d_set_schedule": (_I, [_P, C.POINTER(StepScalars), _I, _I, _I]),
    "ddpm3d_set_sampler": (_I, [_P, _I, C.c_float]),
    "ddpm3d_p_sample_update": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, C.c_int64, _P]),
    "ddpm3d_p_sample": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "ddpm3d_p_sample_t": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "ddpm3d_sample_loop": (_I, [_P, _P, _P, _P, _P, C.c_uint64, _I, _I, _P, _I, _I, _I, _I, _P]),
    "ddpm3d_comm_unique_id": (_I, [_P]),
    "ddpm3d_set_comm": (_I, [_P, _P, _I, _I]),
    "ddpm3d_set_slab": (_I, [_P, _I, _I]),
    "ddpm3d_set_option": (_I, [_P, C.c_char_p, C.c_int64]),
    "ddpm3d_launch_count": (C.c_int64, [_P]),
    "ddpm3d_profile_read": (_I, [_P, C.POINTER(ProfRecord), _I]),
    "ddpm3d_k_conv3d": (_I, [_I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "ddpm3d_k_conv_plan": (_I, [_I] * 12 + [C.POINTER(C.c_int32)]),
    "ddpm3d_k_probe_rowshift": (_I, [_P, _I, _P, _I, _I, _P, _P]),
    "ddpm3d_k_groupnorm": (_I, [_I, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "ddpm3d_k_conv3d_gn": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ddpm3d_k_timestep_embedding": (_I, [_P, _P, _P, _I, _I, _P]),
    "ddpm3d_k_attention": (_I, [_I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ddpm3d_k_attention_window": (_I, [_I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "ddpm3d_k_extract_patch": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "ddpm3d_k_hann_accumulate": (_I, [_P, _P, C.c_double, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "ddpm3d_k_hann_finalize": (_I, [_P, _P, C.c_int64, _P]),
    "ddpm3d_k_welford_update": (_I, [_P, _P, _P, _I, C.c_int64, _P]),
    "ddpm3d_k_welford_merge": (_I, [_P, _P, _I, _P, _P, _I, C.c_int64, _P]),
}


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ddpm3d error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Loads libddpm3d.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(or `python 3d-denoising-diffusion-model_b200/_build.py`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int):
    if code < 0:
        raise NativeError(code, lib().ddpm3d_last_error().decode())
    return code


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
