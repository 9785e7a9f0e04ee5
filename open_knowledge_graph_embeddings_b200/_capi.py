"""ctypes binding of ``libokge_b200.so`` (the C ABI declared in ``include/okge_b200.h``).

This is the only place where Python touches the native library. Tensors are passed as raw device
pointers (``tensor.data_ptr()``) plus sizes; the CUDA stream is torch's current stream. There is
no CPU fallback: if the library is missing, or a call returns a non-zero status, an exception is
raised (``OkgeNativeError``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_uint64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libokge_b200.so")

OKGE_OK = 0
ABI_VERSION = 2
POOL_MODES = {"sum": 0, "mean": 1, "max": 2}
FOLD_COMPLEX_SP, FOLD_COMPLEX_PO, FOLD_DISTMULT = 0, 1, 2

P = c_void_p  # every device pointer
I64 = c_int64
I32 = c_int32
F32 = c_float

# name -> argtypes (restype is int unless listed in _RESTYPES). Mirrors include/okge_b200.h 1:1.
SIGNATURES = {
    "okge_abi_version": [],
    "okge_last_error": [],
    "okge_device_check": [],
    "okge_gather_rows": [P, I64, P, I64, I64, P, I64, P],
    "okge_scatter_add_rows": [P, I64, P, I64, I64, I32, P, I64, P],
    "okge_gather_pool_fwd": [P, I64, P, I32, P, I64, I64, I64, I32, P, I64, P],
    "okge_gather_pool_bwd": [P, I64, P, I64, P, I32, P, I64, I64, I64, I32, P, P],
    "okge_gather_pool_bwd_slots": [P, I64, P, I64, P, I32, P, I64, I64, I64, I32, P, P, P],
    "okge_dropout": [P, I64, F32, c_uint64, c_uint64, P, P],
    "okge_dropout_step": [P, I64, F32, c_uint64, c_uint64, P, P, P],
    "okge_bn_workspace_bytes": [I64, I32, I32],
    "okge_bn_train_fwd": [P, I64, P, I32, I64, I32, P, P, P, P, P, F32, F32, P, I64, P, P, F32, c_uint64, c_uint64, P, P, P],
    "okge_bn_train_bwd": [P, I64, P, I64, P, I32, I64, I32, P, P, P, P, I64, P, P, F32, c_uint64, c_uint64, P, P, P],
    "okge_bn_eval_fwd": [P, I64, I64, I32, P, P, P, P, F32, P, I64, P],
    "okge_bn_col_sums": [P, I64, P, I64, P, P, I64, I32, P, P, P],
    "okge_bn_normalize": [P, I64, I64, I32, P, P, P, P, P, I64, P],
    "okge_bn_normalize_bwd": [P, I64, P, I64, I64, I32, P, P, P, P, P, I64, P],
    "okge_lstm_cell_fwd": [P, I64, P, I64, P, P, P, I64, I64, I32, P, P, P, P, P, P],
    "okge_lstm_cell_bwd": [P, P, P, P, P, I32, P, P, I64, I64, P, P],
    "okge_fold_query": [I32, P, P, I64, I64, P, P],
    "okge_fold_query_rows": [P, P, P, I64, I64, P, P],
    "okge_fold_query_rows_bwd": [P, P, P, P, I64, I64, P, P, P],
    "okge_batch_layout": [P, P, I64, I64, I32, I32, P, P, I32, P, P],
    "okge_fold_query_bwd": [I32, P, P, P, I64, I64, P, P, P],
    "okge_f16_absmax": [P, I64, I64, I64, P, P],
    "okge_f16_quantize": [P, I64, I64, I64, P, F32, P, P, I64, P, P],
    "okge_gemm_f16_nt": [P, I64, I32, P, I64, I32, I64, I64, I64, F32, P, P, P, P, I64, I32, P, P],
    "okge_gemm_tf32_nt": [P, I64, I32, P, I64, I32, I64, I64, I64, F32, P, P, I64, I32, P, P],
    "okge_score_store": [P, P, I64, P, P, I64, I64, I64, I64, P, P, P, I64, P],
    "okge_score_bce": [P, I64, P, I64, I64, I64, I64, P, P, P, P, F32, F32, P, P, P, F32, P],
    "okge_score_bce_rank": [P, P, I64, P, P, I64, I64, I64, I64, I64, P, P, P, P, F32, F32, P, P, P, I32, P, P],
    "okge_score_lse_ws_floats": [I64, I64],
    "okge_score_lse": [P, I64, P, I64, I64, I64, I64, P, P, P, P, P, P, P, P],
    "okge_score_softmax_grad": [P, I64, P, I64, I64, I64, I64, P, P, P, P, P, P, P, F32, P],
    "okge_rank_count": [P, I64, I64, I64, P, P, P, I64, P, P, P, P, P, P],
    "okge_score_rank": [P, P, I64, P, P, I64, I64, I64, I64, P, P, P, P, P, P],
    "okge_rank_true_score": [P, I64, P, P, P, I64, P, P],
    "okge_rank_filter_correct": [P, I64, P, I64, P, P, P, I32, P, P, P],
    "okge_adagrad_dense": [P, P, P, I64, F32, F32, F32, P],
    "okge_adagrad_rows": [P, P, I64, P, I64, P, P, I64, I64, F32, F32, F32, P],
    "okge_gemm_adagrad": [P, I64, I32, P, I64, I32, I64, I64, I64, F32, P, P, P, P, P, I64, P, P, I64, P, I64, P, F32, F32, F32, P],
    "okge_gemm_adagrad_dropout": [P, I64, I32, P, I64, I32, I64, I64, I64, F32, P, P, P, P, P, I64, P, P, I64, P, I64, P, F32, F32,
                                  F32, F32, c_uint64, c_uint64, P, P],
    "okge_f16_mask_dropout": [P, I64, I64, I64, F32, c_uint64, c_uint64, P, P, P, I64, P, P],
    "okge_row_slots_build": [P, I64, I32, P, P],
    "okge_row_slots_accumulate": [P, I64, P, I64, I64, I32, P, P, I64, P],
    "okge_row_slots_clear": [P, I64, I32, P, P],
    "okge_adagrad_slot_rows": [P, P, I64, I64, I64, P, P, I64, F32, F32, F32, P],
    "okge_adagrad_slot_table": [P, P, I64, I64, P, P, F32, F32, F32, P],
    "okge_adam_dense": [P, P, P, P, I64, F32, F32, F32, F32, F32, F32, F32, P],
    "okge_adam_rows": [P, P, P, I64, P, I64, P, P, I64, I64, F32, F32, F32, F32, F32, F32, F32, P],
    "okge_collate_shared": [P, I64, P, P, P, P, I64, I32, I64, I64, I64, I64, c_uint64] + [P] * 15 + [P],
    # host-side (no stream argument, HOST pointers): see dataset.collate_many
    "okge_host_collate_plan": [P, I64, I64, P, I64, P, P],
    "okge_host_collate_fill": [P, I64, I64, P, P, P, P, P, P, P, P],
}
_RESTYPES = {"okge_last_error": c_char_p, "okge_score_lse_ws_floats": c_int64, "okge_bn_workspace_bytes": c_int64}


class OkgeNativeError(RuntimeError):
    """A call into libokge_b200.so failed (or the library could not be loaded)."""


_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the shared library once and declare every prototype. Fails loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OkgeNativeError(
            f"{LIB_PATH} not found: build it with `make -C {os.path.dirname(LIB_PATH)}` or "
            f"`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    if lib.okge_abi_version() != ABI_VERSION:
        raise OkgeNativeError("libokge_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().okge_last_error()
    return msg.decode() if msg else ""


def check(status: int, what: str) -> None:
    if status != OKGE_OK:
        raise OkgeNativeError(f"{what} failed with status {status}: {last_error()}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream. ``torch.cuda.current_stream()`` builds a Stream object per call (~17 us,
    a quarter of the host time of a launch-bound step); the raw getter is a plain C call."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise OkgeNativeError("expected a CUDA tensor; the hot path has no CPU fallback")
    return t.data_ptr()


# Optional observer of every native call: ``hook(name, args, phase)`` with phase "before" / "after", both
# invoked on the launching thread around the asynchronous launch (bench.py records CUDA events there to
# time each kernel on the stream it runs on, and counts launches).
_call_hook = None


def set_call_hook(hook) -> None:
    global _call_hook
    _call_hook = hook


def call(name: str, *args) -> None:
    """Invoke ``name`` on torch's current stream and raise on a non-zero status."""
    lib = load()
    hook = _call_hook
    if hook is not None:
        hook(name, args, "before")
    check(getattr(lib, name)(*args, stream_ptr()), name)
    if hook is not None:
        hook(name, args, "after")
