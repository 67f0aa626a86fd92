"""ctypes binding of libhnm_b200.so (the C ABI in include/hnm_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or the device is
not a B200, every call raises.  torch is used only to own device memory and
streams; the signatures carry raw pointers and sizes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libhnm_b200.so")

_lib: Optional[C.CDLL] = None
ABI_VERSION = 2

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F32 = C.c_float
F64 = C.c_double
SZ = C.c_size_t

_SIGNATURES = {
    "hnm_abi_version": (C.c_int, []),
    "hnm_strerror": (C.c_char_p, [C.c_int]),
    "hnm_check_device": (C.c_int, []),
    "hnm_graph_build_workspace_bytes": (SZ, [I64, I64, C.c_int]),
    "hnm_graph_build": (C.c_int, [P, P, P, I64, I64, P, P, P, P, I32, P, P, P, SZ, P]),
    "hnm_lightgcn_prescale": (C.c_int, [P, P, F32, P, P, I64, I32, P]),
    "hnm_lightgcn_layer": (C.c_int, [P, P, P, P, P, P, P, F32, I64, I32, I64, I64, P, I32, I32, I32, I32, P]),
    "hnm_lightgcn_partial": (C.c_int, [P, P, P, P, P, P, I32, I64, I64, P, I32, I32, I32, I32, P]),
    "hnm_lightgcn_finish": (C.c_int, [P, P, P, F32, P, P, I64, I64, I32, P]),
    "hnm_lightgcn_partial_peer": (C.c_int, [P, P, P, P, P, P, I32, I32, I32, I64, I64, P, I32, I32, I32, P]),
    "hnm_lightgcn_finish_peer": (C.c_int, [P, I32, I32, I32, P, P, F32, P, P, P, I64, I64, I32, P]),
    "hnm_pair_scores": (C.c_int, [P, P, P, P, I64, I32, I64, I64, P, P, P]),
    "hnm_score_all_items": (C.c_int, [P, P, P, I64, I64, I32, P, P]),
    "hnm_topk_exact": (C.c_int, [P, P, P, I64, I64, I64, I32, I32, P, P, I32, P, P, P]),
    "hnm_score_pack_items": (C.c_int, [P, I64, I64, I32, P, P, P, P]),
    "hnm_score_pack_users": (C.c_int, [P, P, I64, I64, I32, P, P, P]),
    "hnm_absmax": (C.c_int, [P, I64, P, I32, P, P]),
    "hnm_column_mean": (C.c_int, [P, I64, I32, P, P, I64, P]),
    "hnm_column_mean_workspace_bytes": (C.c_int64, [I32]),
    "hnm_score_topk_fused": (C.c_int, [P, I64, I64, P, I64, I64, I32, I32, P, I32, P, P, P, P, I64, P]),
    "hnm_exclusion_signature": (C.c_int, [P, P, I64, I64, I64, P, P]),
    "hnm_score_topk_fused_workspace_bytes": (C.c_int64, [I64, I64]),
    "hnm_score_topk_fused_plan": (C.c_int, [I64, I64, P]),
    "hnm_rescore_topk": (C.c_int, [P, P, P, I64, I32, I64, I64, P, I32, P, P, P, P, P, P, P, I32, P, P, P, P]),
    "hnm_merge_topk": (C.c_int, [P, P, I32, I64, I32, P, P, P]),
    "hnm_topk_dense": (C.c_int, [P, I64, I64, P, P, I32, P, P, P]),
    "hnm_ncf_precompute": (C.c_int, [P, I64, I32, P, I32, I32, I32, P, P, P]),
    "hnm_ncf_score_pairs": (C.c_int, [P, P, P, P, P, P, I32, P, F32, P, P, I64, I32, P, P]),
    "hnm_ncf_score_candidates": (C.c_int, [P, P, P, P, P, P, I32, P, F32, P, I64, P, I32, I32, P, P]),
}

_PENDING = set()
EXPORTED = tuple(n for n in _SIGNATURES if n not in _PENDING)


class HnmError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.fn, self.code = fn, code


def load() -> C.CDLL:
    """Load the shared library (no device needed).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m hnm_recommendation_b200.build` "
                "(there is no CPU fallback for the B200 scoring path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            if name in _PENDING:
                continue
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.hnm_abi_version() != ABI_VERSION:
            raise RuntimeError("libhnm_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def strerror(code: int) -> str:
    return load().hnm_strerror(code).decode()


def check(fn: str, code: int) -> None:
    if code != 0:
        raise HnmError(fn, code, strerror(code))


LAUNCHES = 0          # kernels launched through call() since the counter was last reset (bench.py reads it)
_LAUNCHES_PER_CALL = {"hnm_graph_build": 3, "hnm_column_mean": 2}


def call(fn: str, *args) -> None:
    """Call an int-returning entry point and raise HnmError on a non-zero status."""
    global LAUNCHES
    check(fn, getattr(load(), fn)(*args))
    n = _LAUNCHES_PER_CALL.get(fn, 1)
    if fn == "hnm_score_topk_fused" and args[14] > 256:
        n = 2                                     # + merge of the per-slice candidate lists
    if fn == "hnm_lightgcn_partial" and args[10]:
        n = 2 + (1 if args[11] else 0)
    if fn == "hnm_lightgcn_partial_peer" and args[12]:
        n = 2 + (1 if args[13] else 0)
    if fn == "hnm_lightgcn_layer" and args[13]:
        n = 2 + (1 if args[14] else 0)            # cluster pass + whole-CTA pass over the long rows + warp-per-row pass
    LAUNCHES += n


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None stays NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("hnm_b200 kernels need CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise RuntimeError("hnm_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_device() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("hnm_recommendation_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    check("hnm_check_device", load().hnm_check_device())
