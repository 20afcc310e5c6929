"""ctypes binding of libldmb200.so (include/ldmb.h).  No torch types cross this boundary.

The library is the product: if it is missing or cannot be loaded this module raises --
there is no Python/CPU fallback for any operator.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDMB_LIB_PATH") or os.path.join(HERE, "libldmb200.so")     # override: A/B runs against another build of the same ABI

LDMB_MAX_LEVELS = 8
BF16, FP32_VALIDATE = 0, 1
VAE_DECODER, VAE_ENCODER = 0, 1
STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "STATE", 4: "UNSUPPORTED", 5: "KERNEL"}


class UNetConfig(C.Structure):
    _fields_ = [("input_channels", C.c_int32), ("num_levels", C.c_int32), ("stem_size", C.c_int32),
                ("blocks", C.c_int32 * LDMB_MAX_LEVELS), ("channels", C.c_int32 * LDMB_MAX_LEVELS)]


class VaeConfig(C.Structure):
    _fields_ = [("image_channels", C.c_int32), ("latent_channels", C.c_int32), ("num_levels", C.c_int32),
                ("channels", C.c_int32 * LDMB_MAX_LEVELS), ("blocks", C.c_int32 * LDMB_MAX_LEVELS)]


class DdimCoef(C.Structure):
    _fields_ = [("c_eps_in", C.c_float), ("c_div", C.c_float), ("c_x0", C.c_float), ("c_eps_out", C.c_float),
                ("sigma", C.c_float), ("final_step", C.c_int32)]


_H = C.c_void_p      # ldmb_handle*
_P = C.c_void_p      # device / host pointer
_I64P = C.POINTER(C.c_int64)
_I32P = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/ldmb.h declares
SIGNATURES = {
    "ldmb_abi_version": (C.c_int, []),
    "ldmb_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_H)]),
    "ldmb_destroy": (None, [_H]),
    "ldmb_last_error": (C.c_char_p, [_H]),
    "ldmb_precision_of": (C.c_int, [_H]),
    "ldmb_set_force_simt": (C.c_int, [_H, C.c_int]),
    "ldmb_set_use_graphs": (C.c_int, [_H, C.c_int]),
    "ldmb_launch_count": (C.c_int64, [_H]),
    "ldmb_check_device_fault": (C.c_int, [_H, _P]),
    "ldmb_poll_device_fault": (C.c_int, [_H]),
    "ldmb_debug_tc_trace": (C.c_int, [_H, C.c_int, _I64P, C.c_int]),
    "ldmb_unet_forward_per_image": (C.c_int, [_H, _P, _P, C.c_int, C.c_int, C.c_int, _I32P, C.c_int, C.POINTER(_P), _I32P,
                                              C.POINTER(DdimCoef), _P, _P]),
    "ldmb_host_draw_plans": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P, C.c_int, _P, _I64P]),
    "ldmb_unet_precompute_film": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.POINTER(_P), _P]),
    "ldmb_set_deterministic": (C.c_int, [_H, C.c_int]),
    "ldmb_debug_skip_classes": (C.c_int, [_H, C.c_uint32]),
    "ldmb_profile_begin": (C.c_int, [_H]),
    "ldmb_profile_end": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double), _I64P]),
    "ldmb_unet_configure": (C.c_int, [_H, C.POINTER(UNetConfig)]),
    "ldmb_unet_load_param": (C.c_int, [_H, C.c_char_p, _P, _I64P, C.c_int, _P]),
    "ldmb_unet_params_missing": (C.c_int, [_H]),
    "ldmb_unet_reserve": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldmb_unet_set_position_table": (C.c_int, [_H, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_unet_forward": (C.c_int, [_H, _P, _P, C.c_int, C.c_int, C.c_int, _I32P, C.c_int, C.POINTER(_P), _I32P,
                                    C.POINTER(DdimCoef), _P, _P]),
    "ldmb_vae_configure": (C.c_int, [_H, C.c_int, C.POINTER(VaeConfig)]),
    "ldmb_vae_load_param": (C.c_int, [_H, C.c_int, C.c_char_p, _P, _I64P, C.c_int, _P]),
    "ldmb_vae_params_missing": (C.c_int, [_H, C.c_int]),
    "ldmb_vae_reserve": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldmb_vae_decode": (C.c_int, [_H, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_vae_encode": (C.c_int, [_H, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_gemm": (C.c_int, [_H, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_conv3x3": (C.c_int, [_H, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_channelnorm_film": (C.c_int, [_H, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_normconv": (C.c_int, [_H, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_mlp_fused": (C.c_int, [_H, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_mlp_fused_attn": (C.c_int, [_H, _P, _P, _P, _P, _P, _P, C.c_int64, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_grouped_conv3x3": (C.c_int, [_H, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ldmb_window_attention": (C.c_int, [_H, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, _P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libldmb200.so (built in-tree by ldm_image_generator_b200.build) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m ldm_image_generator_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.ldmb_abi_version() != 1:
        raise RuntimeError("libldmb200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib
