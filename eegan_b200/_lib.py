"""ctypes binding of libeegan_b200.so (C ABI: include/eegan_b200.h).

There is no CPU fallback and no other backend: if the shared library is missing, or a
tensor is not a CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeegan_b200.so")

_c_int, _c_float, _c_double, _c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/eegan_b200.h one-to-one
SIGNATURES = {
    "eegan_abi_version": (_c_int, []),
    "eegan_last_error": (ctypes.c_char_p, []),
    "eegan_damsm_pair_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "eegan_damsm_pair_fwd": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                      _p, _p, _c_int, _p, _c_size_t, _p]),
    "eegan_damsm_pair_bwd": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                      _p, _p, _p, _p, _c_size_t, _p]),
    "eegan_damsm_pair_bwd_phased": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                             _p, _p, _p, _c_int, _p, _c_size_t, _p]),
    "eegan_pair_ce_fwd": (_c_int, [_p, _c_float, _p, _p, _c_int, _p, _p, _p, _p]),
    "eegan_pair_ce_bwd": (_c_int, [_p, _p, _p, _p, _c_float, _c_int, _p, _p]),
    "eegan_sent_scores_fwd": (_c_int, [_p, _p, _c_int, _c_int, _c_float, _c_float, _p, _p, _p]),
    "eegan_sent_scores_bwd": (_c_int, [_p, _p, _p, _p, _c_int, _c_int, _c_float, _c_float, _p, _p, _p]),
    "eegan_gag_fwd": (_c_int, [_p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _p, _p]),
    "eegan_gag_bwd": (_c_int, [_p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p]),
    "eegan_gag_bwd_workspace_bytes": (_c_size_t, [_c_int] * 4),
    "eegan_gag_bwd_ws": (_c_int, [_p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p, _c_size_t, _p]),
    "eegan_syncbn_stats": (_c_int, [_p, _c_int, _c_int, _c_int, _p, _p]),
    "eegan_syncbn_stats_counted": (_c_int, [_p, _c_int, _c_int, _c_int, _p, _p]),
    "eegan_syncbn_fwd_fused": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int, _c_float, _c_float, _p, _p, _p, _p, _p]),
    "eegan_syncbn_bwd_fused": (_c_int, [_p, _p, _p, _p, _c_int, _c_int, _c_int, _c_float, _p, _p, _p]),
    "eegan_ssa_fwd_fused": (_c_int, [_p, _p, _p, _p, _c_int, _c_int, _c_int, _c_float, _c_float, _p, _p, _p, _p, _p]),
    "eegan_syncbn_finalize": (_c_int, [_p, _c_int, _c_double, _p, _c_float, _c_float, _c_int, _p, _p, _p, _p, _p]),
    "eegan_syncbn_apply": (_c_int, [_p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _p, _p]),
    "eegan_syncbn_bwd_reduce": (_c_int, [_p, _p, _p, _p, _c_int, _c_int, _c_int, _p, _p]),
    "eegan_syncbn_bwd_apply": (_c_int, [_p, _p, _p, _p, _p, _p, _c_double, _p, _c_float, _c_int, _c_int, _c_int, _c_int,
                                        _p, _p]),
    "eegan_func_attention_workspace_bytes": (_c_size_t, [_c_int] * 4),
    "eegan_func_attention_fwd": (_c_int, [_p, _p, _c_int, _c_int, _c_int, _c_int, _c_float, _p, _p, _p, _c_size_t, _p]),
    "eegan_func_attention_bwd": (_c_int, [_p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_float, _p, _p,
                                          _p, _c_size_t, _p]),
    "eegan_cosine_rows_fwd": (_c_int, [_p, _p, ctypes.c_longlong, _c_int, _c_float, _p, _p, _p]),
    "eegan_cosine_rows_bwd": (_c_int, [_p, _p, _p, _p, _p, ctypes.c_longlong, _c_int, _c_float, _p, _p, _p]),
    "eegan_gemm_tf32x3": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int] + [ctypes.c_longlong] * 6 + [_c_int, _c_int, _p]),
    "eegan_gemm_f16x3": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int] + [ctypes.c_longlong] * 6 + [_c_int, _c_float, _c_float, _c_int, _p, _c_size_t, _p]),
    "eegan_conv1x1_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "eegan_conv1x1_fwd": (_c_int, [_p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _c_size_t, _p]),
    "eegan_conv1x1_bwd": (_c_int, [_p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _c_size_t, _p]),
    "eegan_attr_enhance_fwd": (_c_int, [_p] * 8 + [_c_int, _c_int, _c_int, _c_float, _p, _p, _p, _p, _p]),
    "eegan_attr_enhance_bwd": (_c_int, [_p] * 8 + [_c_int, _c_int, _c_int, _c_float] + [_p] * 11),
    "eegan_attr_enhance_workspace_bytes": (_c_size_t, [_c_int] * 3),
    "eegan_rprecision": (_c_int, [_p, _p, _c_int, _c_int, _c_int, _c_float, _p, _p, _p, _p]),
    "eegan_ssa_apply": (_c_int, [_p] * 6 + [_c_int] * 3 + [_p, _p]),
    "eegan_ssa_bwd_reduce": (_c_int, [_p] * 7 + [_c_int] * 3 + [_p] * 5),
    "eegan_ssa_bwd_apply": (_c_int, [_p] * 7 + [_c_double, _p, _c_float, _c_int, _c_int, _c_int, _c_int, _p, _p]),
    "eegan_set_gag_engine": (_c_int, [_c_int]),
    "eegan_get_gag_engine": (_c_int, []),
    "eegan_set_gag_bwd_engine": (_c_int, [_c_int]),
    "eegan_get_gag_bwd_engine": (_c_int, []),
    "eegan_set_contraction_engine": (_c_int, [_c_int]),
    "eegan_get_contraction_engine": (_c_int, []),
    "eegan_profile_enable": (_c_int, [_c_int]),
    "eegan_profile_nstages": (_c_int, []),
    "eegan_profile_stage_name": (ctypes.c_char_p, [_c_int]),
    "eegan_profile_collect": (_c_int, [_p, _p]),
}

_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        "eegan_b200: %s is missing — build it with `make -C eegan_b200/csrc` "
                        "(or __graft_entry__.build()).  There is no CPU or PyTorch fallback." % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)  # AttributeError if the .so is stale
                    fn.restype, fn.argtypes = res, args
                if handle.eegan_abi_version() != 1:
                    raise RuntimeError("eegan_b200: ABI version mismatch")
                _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().eegan_last_error().decode("utf-8", "replace")
        raise RuntimeError("eegan_b200 %s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("eegan_b200: expected CUDA tensors (this library has no CPU path); got device %s" % t.device)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """Contiguous fp32 view/copy (the reference's inputs are contiguous fp32 already)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
