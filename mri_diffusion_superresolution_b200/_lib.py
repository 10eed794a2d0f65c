"""ctypes binding of libmrisr_b200.so (C ABI declared in include/mrisr_b200.h).

There is no CPU path: if the shared library is missing the import of any product module that needs it raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmrisr_b200.so")

ABI_VERSION = 6
ACT_NONE, ACT_RELU, ACT_SILU, ACT_GEGLU = 0, 1, 2, 3
E_INVALID, E_UNSUPPORTED, E_CUDA = -1, -2, -3


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("n_store", C.c_int32),
        ("k1", C.c_int32), ("k2", C.c_int32), ("taps", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32),
        ("a1", C.c_void_p), ("lda1", C.c_int64),
        ("a2", C.c_void_p), ("lda2", C.c_int64),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p), ("rowvec_stride", C.c_int64),
        ("rows_per_batch", C.c_int32), ("act", C.c_int32),
        ("res1", C.c_void_p), ("ldr1", C.c_int64),
        ("res2", C.c_void_p), ("ldr2", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out_fp32", C.c_int32), ("reserved", C.c_int32),
        ("conv_stride", C.c_int32), ("conv_pad_mode", C.c_int32),
        ("f16_flags", C.c_int32), ("reserved3", C.c_int32),
        ("gn_stats", C.c_void_p), ("ld_stats", C.c_int64),
        ("lora_a", C.c_void_p), ("lora_n", C.c_int32), ("reserved4", C.c_int32),
        ("lora_t_out", C.c_void_p),
        ("splitk_ws", C.c_void_p), ("splitk_ws_floats", C.c_int64),
    ]


class AdamDesc(C.Structure):
    """mrisr_adam_desc (include/mrisr_b200.h)."""
    _fields_ = [
        ("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
        ("g", C.c_void_p), ("g_sr", C.c_int64), ("g_sc", C.c_int64), ("g_scale", C.c_float),
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("d1", C.c_void_p), ("d1_sr", C.c_int64), ("d1_sc", C.c_int64), ("d1_scale", C.c_float), ("d1_f16", C.c_int32),
        ("d2", C.c_void_p), ("d2_sr", C.c_int64), ("d2_sc", C.c_int64), ("d2_scale", C.c_float), ("d2_f16", C.c_int32),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/mrisr_b200.h declares.
PROTOTYPES = {
    "mrisr_abi_version": (_I, []),
    "mrisr_last_error": (C.c_char_p, []),
    "mrisr_device_info": (_I, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mrisr_sched_step": (_I, [_P, _P, _P, _P, _P, _L, _P, _P]),
    "mrisr_res_shift": (_I, [_P, _P, _P, _P, _L, _I, _P, _I, _P, _I, _P]),
    "mrisr_sched_step_indexed": (_I, [_P, _P, _P, _P, _L, _P, _L, _P, _P, _I, _P]),
    "mrisr_select_row": (_I, [_P, _P, _I, _L, _P, _I, _P]),
    "mrisr_advance_index": (_I, [_P, _P]),
    "mrisr_timestep_embedding": (_I, [_P, _P, _I, _I, _P]),
    "mrisr_sinusoidal_embedding": (_I, [_P, _P, _I, _I, _I, _P]),
    "mrisr_groupnorm_workspace_floats": (_L, [_I, _I]),
    "mrisr_groupnorm": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _P, _P, _F, _I, _P, _P, _I, _P]),
    "mrisr_groupnorm_apply_stats": (_I, [_P, _L, _I, _P, _L, _I, _L, _P, _L, _I, _P, _L, _I, _L, _I, _I, _I, _P, _P, _F, _I, _P, _P, _I, _P]),
    "mrisr_layernorm": (_I, [_P, _L, _P, _P, _F, _P, _L, _I, _I, _I, _P]),
    "mrisr_gemm": (_I, [C.POINTER(GemmArgs), _P]),
    "mrisr_gemm_block_n": (_I, [_I, _I]),
    "mrisr_attention": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _P]),
    "mrisr_upsample2x": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mrisr_im2col3x3s2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mrisr_im2col_first": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mrisr_pixel_unshuffle_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mrisr_avgpool2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mrisr_add": (_I, [_P, _P, _P, _L, _I, _P]),
    "mrisr_transpose": (_I, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "mrisr_cast": (_I, [_P, _I, _P, _I, _L, _P]),
    "mrisr_bilinear_resize": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mrisr_to_uint8_vis": (_I, [_P, _P, _I, _I, _I, _P]),
    "mrisr_softmax_rows": (_I, [_P, _L, _P, _L, _I, _I, _F, _P]),
    "mrisr_channel_mix": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mrisr_gaussian_sample": (_I, [_P, _P, _P, _I, _I, _I, _F, _P]),
    "mrisr_eval_metrics_workspace_floats": (_L, [_I, _I, _I]),
    "mrisr_eval_metrics": (_I, [_P, _P, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P]),
    "mrisr_groupnorm_backward": (_I, [_P, _L, _I, _P, _L, _I, _P, _I, _I, _I, _P, _P, _F, _I, _P, _L, _P, _L, _P, _I, _P]),
    "mrisr_layernorm_backward": (_I, [_P, _L, _I, _P, _P, _F, _P, _P, _I, _I, _P]),
    "mrisr_geglu_forward": (_I, [_P, _P, _L, _I, _P]),
    "mrisr_geglu_backward": (_I, [_P, _P, _P, _L, _I, _P]),
    "mrisr_zero_insert2x": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mrisr_sumpool2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mrisr_mse_grad": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "mrisr_xty64_workspace_floats": (_L, [_I, _I]),
    "mrisr_xty64": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _F, _P, _P, _P]),
    "mrisr_gemm_splitk_workspace_floats": (_L, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "mrisr_attention_backward_workspace": (_L, [_I, _I, _I, _I, _I]),
    "mrisr_attention_backward": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _P]),
    "mrisr_attention_exports_lse": (_I, [_I, _I, _L]),
    "mrisr_attention_lse": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _P]),
    "mrisr_grad_sqnorm": (_I, [_P, _I, _F, _P, _P, _P]),
    "mrisr_adamw": (_I, [_P, _I, _P, _P, _F, _F, _F, _F, _P, _P]),
    "mrisr_slice_volume": (_I, [_P, _I, _I, _I, _I, _F, _F, _F, _P, _I, _I, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m mri_diffusion_superresolution_b200._build` "
                "(or __graft_entry__.build()).  This package has no CPU or PyTorch fallback path.")
        import torch  # noqa: F401  -- maps libcudart.so.12 (the library links the shared runtime) before the dlopen
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mrisr_abi_version() != ABI_VERSION:
            raise RuntimeError("libmrisr_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


# number of CUDA kernels launched through the C ABI by this process (bench.py reports it as gpu_launches)
LAUNCHES = [0]


def check(code: int, what: str, kernels: int = 1) -> None:
    """Map a negative ABI return code to the Python exceptions the reference's callers would see
    (ValueError for bad arguments, cf. src/adapters/modules.py:16,29; RuntimeError otherwise)."""
    if code == 0:
        LAUNCHES[0] += kernels
        return
    msg = load().mrisr_last_error().decode("utf-8", "replace")
    if code in (E_INVALID, E_UNSUPPORTED):
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (code {code})")
