"""ctypes binding of libvdetr_b200.so (C ABI declared in include/vdetr_b200.h).

The reference binds its native ops with pybind11 (third_party/pointnet2/_ext_src/src/bindings.cpp:9-21);
here the boundary is a plain C ABI so that it can be bound from anything.  Only raw
device pointers, sizes and the current CUDA stream cross the boundary.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VDETR_B200_LIB") or os.path.join(_HERE, "lib", "libvdetr_b200.so")   # (override: developer experiments)
_lib = None

c_int, c_size_t, c_void_p, c_float = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float


class XattnShape(ctypes.Structure):
    _fields_ = [("B", c_int), ("nQ", c_int), ("nK", c_int), ("H", c_int), ("hd", c_int), ("grid_n", c_int),
                ("log_scale", c_float), ("max_value", c_float), ("rotate", c_int), ("kv_heads", c_int),
                ("has_bias", c_int)]


PEER_MAX_WORLD, PEER_CHANNELS = 8, 4


class PeerCtx(ctypes.Structure):
    _fields_ = [("flags", c_void_p * PEER_MAX_WORLD), ("slots", c_void_p * PEER_MAX_WORLD), ("rank", c_int), ("world", c_int),
                ("epoch", c_void_p), ("cap", c_int)]


_SIGS = {
    "vdetr_version": (ctypes.c_char_p, []),
    "vdetr_error_string": (ctypes.c_char_p, [c_int]),
    "vdetr_pn2_fps_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vdetr_pn2_fps": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vdetr_pn2_fps_ragged_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vdetr_pn2_fps_ragged": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vdetr_pn2_gather": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vdetr_pn2_gather_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vdetr_pn2_ball_query": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "vdetr_pn2_group": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vdetr_pn2_group_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vdetr_xattn_fwd_workspace_bytes": (c_size_t, [ctypes.POINTER(XattnShape), c_int]),
    "vdetr_xattn_bias_save_bytes": (c_size_t, [ctypes.POINTER(XattnShape), c_int]),
    "vdetr_xattn_fwd": (c_int, [ctypes.POINTER(XattnShape)] + [c_void_p] * 10 + [c_float, c_void_p] +
                        [c_void_p, c_size_t, c_int, c_void_p]),
    "vdetr_xattn_bwd_workspace_bytes": (c_size_t, [ctypes.POINTER(XattnShape), c_int, c_int]),
    "vdetr_xattn_bwd": (c_int, [ctypes.POINTER(XattnShape)] + [c_void_p] * 11 + [c_float, c_void_p] + [c_void_p] * 4 +
                        [c_void_p, c_size_t, c_int, c_void_p]),
    "vdetr_rpe_bias": (c_int, [ctypes.POINTER(XattnShape)] + [c_void_p] * 5 + [c_void_p]),
    "vdetr_launch_count": (ctypes.c_ulonglong, [c_int]),
    "vdetr_timing_enable": (c_int, [c_int]),
    "vdetr_timing_read": (c_int, [ctypes.POINTER(c_float), ctypes.POINTER(c_int)]),
    "vdetr_layernorm_supported": (c_int, [c_int]),
    "vdetr_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vdetr_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "vdetr_reduce_workspace_floats": (c_size_t, [c_int]),
    "vdetr_colsum_workspace_floats": (c_size_t, [c_int]),
    "vdetr_adamw_flat": (c_int, [c_void_p] * 4 + [ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p, c_void_p] + [c_float] * 5 + [c_void_p]),
    "vdetr_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), ctypes.c_char_p]),
    "vdetr_peer_open": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "vdetr_peer_close": (c_int, [c_void_p]),
    "vdetr_peer_free": (c_int, [c_void_p]),
    "vdetr_peer_error": (c_int, [ctypes.POINTER(c_int)]),
    "vdetr_peer_barrier": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vdetr_peer_norm_bytes": (c_size_t, []),
    "vdetr_adamw_flat_peer": (c_int, [c_void_p] * 4 + [c_int, c_int] + [c_void_p] * 4 + [ctypes.c_longlong] * 4 + [c_void_p, c_void_p] +
                              [c_float, c_float, c_void_p] + [c_float] * 4 + [c_void_p]),
    "vdetr_peer_bn_slot_floats": (c_size_t, [c_int]),
    "vdetr_peer_bn_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  ctypes.c_longlong, c_int, c_void_p, c_void_p]),
    "vdetr_peer_bn_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "vdetr_bn_sync_set": (c_int, [c_void_p]),
    "vdetr_colsum": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vdetr_bn_relu_supported": (c_int, [c_int]),
    "vdetr_bn_relu_train_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_longlong, ctypes.c_longlong,
                                        c_float, c_float, c_float, c_void_p] + [c_void_p] * 7),
    "vdetr_bn_relu_train_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, ctypes.c_longlong, ctypes.c_longlong, c_float] +
                                [c_void_p] * 5),
    "vdetr_box_decode_fwd": (c_int, [c_void_p] * 6 + [c_int, c_int] + [c_void_p] * 8 + [c_void_p]),
    "vdetr_box_decode_bwd": (c_int, [c_void_p] * 9 + [c_int, c_int] + [c_void_p] * 2 + [c_void_p]),
    "vdetr_lsap": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vdetr_debug_dt_clocks": (c_int, [ctypes.POINTER(ctypes.c_ulonglong)]),
    "vdetr_rpe_dtables_workspace_bytes": (c_size_t, [ctypes.POINTER(XattnShape)]),
    "vdetr_rpe_dtables": (c_int, [ctypes.POINTER(XattnShape)] + [c_void_p] * 5 + [c_void_p, c_size_t, c_void_p]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load the shared library (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python v-detr_b200/build.py` (or __graft_entry__.build()). "
                               "vdetr_b200 has no CPU / PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int):
    if code != 0:
        raise RuntimeError(f"{lib().vdetr_error_string(code).decode()} (code {code})")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda(name, t, dtype=None):
    """Same contract as the reference's CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT
    (third_party/pointnet2/_ext_src/include/utils.h:8-28): violations raise RuntimeError."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (CPU not supported)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {dtype} tensor")
    return t
