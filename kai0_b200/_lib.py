"""ctypes binding of libpi05.so (the C-ABI in include/pi05.h).

There is deliberately NO fallback: if the shared library is missing or fails to load, importing this module's
`lib()` raises.  Nothing in the product path computes on the CPU or through torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpi05.so")

_lock = threading.Lock()
_lib = None


class GemmaCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "depth", "mlp_dim", "num_heads", "num_kv_heads", "head_dim")]


class Config(C.Structure):
    _fields_ = [
        ("paligemma", GemmaCfg),
        ("expert", GemmaCfg),
        ("vit_width", C.c_int32),
        ("vit_depth", C.c_int32),
        ("vit_mlp_dim", C.c_int32),
        ("vit_heads", C.c_int32),
        ("vit_patch", C.c_int32),
        ("image_size", C.c_int32),
        ("vocab_size", C.c_int32),
        ("action_dim", C.c_int32),
        ("action_horizon", C.c_int32),
        ("max_token_len", C.c_int32),
        ("num_images", C.c_int32),
        ("max_batch", C.c_int32),
        ("train", C.c_int32),
        ("value_head", C.c_int32),
        ("rtc", C.c_int32),
    ]


class Param(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("dtype", C.c_int32),
        ("numel", C.c_int64),
        ("data", C.c_void_p),
        ("grad", C.c_void_p),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("images", C.c_void_p),
        ("image_masks", C.c_void_p),
        ("tokens", C.c_void_p),
        ("token_mask", C.c_void_p),
        ("patch_rows", C.c_void_p),
        ("token_len", C.c_int32),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("batch", C.c_int32),
        ("batch_inner", C.c_int32),
        ("A", C.c_void_p),
        ("B", C.c_void_p),
        ("a_major", C.c_int32),
        ("b_major", C.c_int32),
        ("lda", C.c_int64),
        ("ldb", C.c_int64),
        ("a_batch_stride", C.c_int64),
        ("b_batch_stride", C.c_int64),
        ("a_batch_stride1", C.c_int64),
        ("b_batch_stride1", C.c_int64),
        ("epilogue", C.c_int32),
        ("D", C.c_void_p),
        ("ldd", C.c_int64),
        ("d_batch_stride", C.c_int64),
        ("d_batch_stride1", C.c_int64),
        ("D2", C.c_void_p),
        ("ldd2", C.c_int64),
        ("d2_batch_stride", C.c_int64),
        ("d2_batch_stride1", C.c_int64),
        ("bias", C.c_void_p),
        ("res", C.c_void_p),
        ("ldres", C.c_int64),
        ("res_batch_stride", C.c_int64),
        ("res_batch_stride1", C.c_int64),
        ("gate", C.c_void_p),
        ("gate_rows", C.c_int32),
        ("ldgate", C.c_int64),
        ("scale", C.c_float),
        ("accumulate", C.c_int32),
        ("block_n", C.c_int32),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
    ]


# Epilogue codes (csrc/gemm.h)
EPI_STORE, EPI_SCALE, EPI_BIAS, EPI_BIAS_GELU, EPI_RES, EPI_GEGLU, EPI_F32 = range(7)

# Every symbol include/pi05.h declares; tests check the library exports all of them.
EXPORTS = (
    "pi05_abi_version",
    "pi05_last_error",
    "pi05_workspace_bytes",
    "pi05_create",
    "pi05_destroy",
    "pi05_bind_params",
    "pi05_params_updated",
    "pi05_forward",
    "pi05_backward",
    "pi05_set_taps",
    "pi05_prefill",
    "pi05_denoise",
    "pi05_denoise_rtc",
    "pi05_forward_advantage",
    "pi05_forward_value",
    "pi05_get_tap",
    "pi05_preprocess_scratch_floats",
    "pi05_preprocess_image",
    "pi05_patch_row_kp",
    "pi05_preprocess_patches",
    "pi05_debug_profile_layer",
    "pi05_debug_set_pdl",
    "pi05_gemm_bf16",
    "pi05_fused_clip_adamw",
    "pi05_fused_clip_adamw_scaled",
    "pi05_set_grad_exchange",
    "pi05_allreduce_grads",
    "pi05_grad_exchange_stats",
    "pi05_nccl_unique_id",
    "pi05_nccl_comm_create",
    "pi05_nccl_comm_destroy",
    "pi05_launch_count",
    "pi05_gemm_profile_enable",
    "pi05_gemm_profile_report",
)


def lib() -> C.CDLL:
    """Load libpi05.so once.  Raises RuntimeError (never falls back) if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or torch fallback for the pi0.5 engine)"
            )
        l = C.CDLL(LIB_PATH)
        l.pi05_abi_version.restype = C.c_int
        l.pi05_last_error.restype = C.c_char_p
        l.pi05_gemm_bf16.restype = C.c_int
        l.pi05_gemm_bf16.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        if hasattr(l, "pi05_create"):
            l.pi05_workspace_bytes.restype = C.c_size_t
            l.pi05_workspace_bytes.argtypes = [C.POINTER(Config)]
            l.pi05_create.restype = C.c_int
            l.pi05_create.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
            l.pi05_destroy.restype = None
            l.pi05_destroy.argtypes = [C.c_void_p]
            l.pi05_bind_params.restype = C.c_int
            l.pi05_bind_params.argtypes = [C.c_void_p, C.POINTER(Param), C.c_int]
            l.pi05_params_updated.restype = C.c_int
            l.pi05_params_updated.argtypes = [C.c_void_p, C.c_void_p]
            l.pi05_forward.restype = C.c_int
            l.pi05_forward.argtypes = [
                C.c_void_p,
                C.POINTER(Batch),
                C.c_void_p,
                C.c_void_p,
                C.c_void_p,
                C.c_void_p,
                C.c_void_p,
            ]
            l.pi05_set_taps.restype = C.c_int
            l.pi05_set_taps.argtypes = [C.c_void_p, C.c_int]
            l.pi05_backward.restype = C.c_int
            l.pi05_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            l.pi05_prefill.restype = C.c_int
            l.pi05_prefill.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p]
            l.pi05_denoise.restype = C.c_int
            l.pi05_denoise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            l.pi05_denoise_rtc.restype = C.c_int
            l.pi05_denoise_rtc.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
            l.pi05_forward_value.restype = C.c_int
            l.pi05_forward_value.argtypes = [C.c_void_p, C.POINTER(Batch)] + [C.c_void_p] * 4
            l.pi05_forward_advantage.restype = C.c_int
            l.pi05_forward_advantage.argtypes = (
                [C.c_void_p, C.POINTER(Batch)] + [C.c_void_p] * 4 + [C.c_float, C.c_float] + [C.c_void_p] * 3
            )
            l.pi05_get_tap.restype = C.c_int
            l.pi05_get_tap.argtypes = [
                C.c_void_p,
                C.c_char_p,
                C.c_void_p,
                C.POINTER(C.c_int64),
                C.POINTER(C.c_int32),
                C.c_void_p,
            ]
        if hasattr(l, "pi05_preprocess_image"):
            l.pi05_preprocess_scratch_floats.restype = C.c_size_t
            l.pi05_preprocess_scratch_floats.argtypes = [C.c_int32, C.c_int32]
            l.pi05_preprocess_image.restype = C.c_int
            l.pi05_preprocess_image.argtypes = [C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p] * 4
            l.pi05_patch_row_kp.restype = C.c_int32
            l.pi05_patch_row_kp.argtypes = [C.c_int32]
            l.pi05_preprocess_patches.restype = C.c_int
            l.pi05_preprocess_patches.argtypes = [C.c_void_p] + [C.c_int32] * 9 + [C.c_void_p] * 4
        if hasattr(l, "pi05_fused_clip_adamw"):
            l.pi05_fused_clip_adamw.restype = C.c_int
            l.pi05_fused_clip_adamw.argtypes = (
                [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5
                + [C.c_int64, C.c_float, C.c_void_p, C.c_void_p]
            )
            l.pi05_fused_clip_adamw_scaled.restype = C.c_int
            l.pi05_fused_clip_adamw_scaled.argtypes = (
                [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5
                + [C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
            )
        if hasattr(l, "pi05_set_grad_exchange"):
            l.pi05_set_grad_exchange.restype = C.c_int
            l.pi05_set_grad_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
            l.pi05_allreduce_grads.restype = C.c_int
            l.pi05_allreduce_grads.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
            l.pi05_grad_exchange_stats.restype = C.c_int
            l.pi05_grad_exchange_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
            l.pi05_nccl_unique_id.restype = C.c_int
            l.pi05_nccl_unique_id.argtypes = [C.c_void_p]
            l.pi05_nccl_comm_create.restype = C.c_int
            l.pi05_nccl_comm_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
            l.pi05_nccl_comm_destroy.restype = C.c_int
            l.pi05_nccl_comm_destroy.argtypes = [C.c_void_p]
        _lib = l
        return _lib


def last_error() -> str:
    return (lib().pi05_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
