"""Tensor-level entry points of the native hot path (thin wrappers over ``_capi``).

Each function validates dtype/layout, allocates the output with torch (device memory and streams
are torch's job here), and calls exactly one C-ABI entry point of ``include/okge_b200.h`` — except
``gemm_nt`` helpers that also pick a split-K factor. Nothing in this module computes on the CPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _capi
from ._capi import FOLD_COMPLEX_PO, FOLD_COMPLEX_SP, FOLD_DISTMULT, POOL_MODES, call, ptr

__all__ = [
    "gather_rows", "scatter_add_rows", "gather_pool_fwd", "gather_pool_bwd", "gather_pool_bwd_slots", "adagrad_slot_table", "dropout", "fold_query",
    "fold_query_bwd", "fold_query_rows", "fold_query_rows_bwd", "gemm_nt", "score_store", "score_bce", "score_bce_rank", "score_lse", "score_softmax_grad",
    "rank_count", "score_rank", "rank_true_score", "rank_filter_correct",
    "adagrad_dense", "adagrad_rows", "adam_dense", "adam_rows", "gemm_adagrad", "mask_dropout_f16", "batch_layout", "row_slots_build",
    "row_slots_accumulate", "row_slots_clear", "adagrad_slot_rows", "bn_train_fwd", "bn_train_bwd", "bn_eval_fwd", "bn_col_sums", "bn_normalize", "bn_normalize_bwd", "lstm_cell_fwd", "lstm_cell_bwd", "pad4", "pad8", "Panels", "MNPanels", "ColMajor",
    "F16Operand", "quantize", "as_f16", "gather_rows_f16", "sm_count", "DS_SCALE_BCE", "DS_SCALE_KL", "TF32_RAW_OPERAND_SCALE",
    "FOLD_COMPLEX_SP", "FOLD_COMPLEX_PO", "FOLD_DISTMULT",
]

# okge_gemm_tf32_nt (fp32 operands outside the 1-vs-all path): tcgen05 truncates raw fp32 operands to TF32;
# 1 / (1 - 2^-11 / (2 ln 2)) centres the error (okge_common.cuh)
TF32_RAW_OPERAND_SCALE = 1.0 / (1.0 - 0.00035221)


def pad4(n: int) -> int:
    """Leading dimensions feeding the tensor-core path must be multiples of 4 floats (16 B)."""
    return (int(n) + 3) // 4 * 4


def pad8(n: int) -> int:
    """Row pitches of fp16 operands are multiples of 8 halves (16 B)."""
    return (int(n) + 7) // 8 * 8


PANEL = 64          # fp16 elements per 128-byte panel line (okge_b200.h, OKGE_K_PANELS)
DS_SCALE_BCE = 4096.0   # dS = sigmoid(s) - y lies in (-1, 1): stored as fp16(4096 dS)
DS_SCALE_KL = 64.0      # softmax gradient: row_weight * p - y with row weights (positives per row) up to ~1000


class F16Operand:
    """An fp16 operand of the tensor-core contractions: the logical fp32 matrix ``x`` [rows, K] as ``hi = fp16(x * scale)``
    (row-major, row pitch padded to 16 B) with ONE power-of-two scale whose inverse lives in device memory
    (``inv_scale`` [1] fp32; None = 1). ``data`` is [planes, rows, ld]: plane 0 = hi, an optional plane 1 = lo =
    fp16(x * scale - hi) (split precision, evaluation). Produced by :func:`quantize` / the fused Adagrad epilogue."""

    def __init__(self, data: torch.Tensor, rows: int, k: int, inv_scale: Optional[torch.Tensor]):
        assert data.dtype == torch.float16 and data.dim() == 3
        self.data, self.rows, self.k, self.inv_scale = data, int(rows), int(k), inv_scale
        self.shape = (self.rows, self.k)
        self.device = data.device

    @property
    def hi(self) -> torch.Tensor:
        return self.data[0]

    @property
    def lo(self) -> Optional[torch.Tensor]:
        return self.data[1] if self.data.size(0) > 1 else None

    @property
    def ld(self) -> int:
        return self.data.stride(1)

    def size(self, dim: int) -> int:
        return self.shape[dim]

    def without_lo(self) -> "F16Operand":
        return self if self.data.size(0) == 1 else F16Operand(self.data[:1], self.rows, self.k, self.inv_scale)

    def row_slice(self, start: int, stop: Optional[int] = None) -> "F16Operand":
        """Rows [start, stop) over the same storage and scale."""
        stop = self.rows if stop is None else stop
        return F16Operand(self.data[:, start:stop], stop - start, self.k, self.inv_scale)

    def dense(self) -> torch.Tensor:
        """The fp32 matrix this operand represents (tests / debugging)."""
        x = self.data[0, :, : self.k].float()
        if self.data.size(0) > 1:
            x = x + self.data[1, :, : self.k].float()
        return x * (self.inv_scale if self.inv_scale is not None else 1.0)


class Panels:
    """A logical [rows, K] fp16 GEMM operand stored as K-panels ``[ceil(K/64), rows, 64]`` (OKGE_K_PANELS in
    okge_b200.h): every TMA box of the tensor-core kernel is one contiguous block of memory. ``scale``: the stored
    values are fp16(scale * x) (a host-known power of two). Produced by the loss epilogues (dS); consumed by
    ``gemm_nt`` / ``gemm_adagrad``."""

    def __init__(self, data: torch.Tensor, rows: int, k: int, scale: float = 1.0):
        self.data, self.rows, self.k, self.scale = data, int(rows), int(k), float(scale)
        self.shape = (self.rows, self.k)
        self.device = data.device

    @staticmethod
    def empty(rows: int, k: int, device, scale: float = 1.0) -> "Panels":
        return Panels(torch.empty(((k + PANEL - 1) // PANEL, rows, PANEL), dtype=torch.float16, device=device), rows, k, scale)

    @staticmethod
    def from_dense(x: torch.Tensor, scale: float = 1.0) -> "Panels":
        """Panels of an fp32 matrix (tests; the product path gets its panels from the loss epilogues)."""
        rows, k = x.shape
        buf = torch.zeros((rows, (k + PANEL - 1) // PANEL * PANEL), dtype=torch.float32, device=x.device)
        buf[:, :k] = x * scale
        return Panels(buf.reshape(rows, -1, PANEL).permute(1, 0, 2).contiguous().to(torch.float16), rows, k, scale)

    def size(self, dim: int) -> int:
        return self.shape[dim]

    def dense(self) -> torch.Tensor:
        """The logical fp32 [rows, K] matrix (tests / debugging)."""
        return (self.data.permute(1, 0, 2).reshape(self.rows, -1)[:, : self.k].float() / self.scale).contiguous()

    @property
    def T(self) -> "MNPanels":
        """The transposed operand [K, rows] over the SAME storage (OKGE_MN_PANELS): no data movement."""
        return MNPanels(self.data, self.k, self.rows, self.scale)


class MNPanels:
    """A logical [rows, K] operand stored as ``[ceil(rows/64), K, 64]`` (OKGE_MN_PANELS): the K-panel storage of the
    transposed matrix, read MN-major by the tensor core. ``dS.T`` is how dE = dS^T Q consumes the loss gradient."""

    def __init__(self, data: torch.Tensor, rows: int, k: int, scale: float = 1.0):
        self.data, self.rows, self.k, self.scale = data, int(rows), int(k), float(scale)
        self.shape = (self.rows, self.k)
        self.device = data.device

    def size(self, dim: int) -> int:
        return self.shape[dim]


class ColMajor:
    """``ColMajor(x)`` presents a row-major ``x[K, rows]`` (fp32 tensor or :class:`F16Operand`) as the logical operand
    ``x^T [rows, K]`` (OKGE_COL_MAJOR) without a transpose pass: E[N, D] enters dQ = dS E and Q[B, D] enters
    dE = dS^T Q this way."""

    def __init__(self, x):
        self.x = x
        self.shape = (x.size(1), x.size(0))
        self.device = x.device


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise _capi.OkgeNativeError(f"{name} must be a CUDA tensor; the hot path has no CPU fallback")
    return t


def _i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.int32:
        raise TypeError(f"{name} must be int32, got {t.dtype}")
    if not t.is_cuda:
        raise _capi.OkgeNativeError(f"{name} must be a CUDA tensor; the hot path has no CPU fallback")
    return t.contiguous()


def _rowmajor(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D fp32 with unit inner stride (an outer stride / leading dimension is allowed)."""
    _f32(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D, got shape {tuple(t.shape)}")
    if t.size(1) > 1 and t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), t.stride(0))


# ---------------------------------------------------------------------------------------------
# embeddings
# ---------------------------------------------------------------------------------------------

def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    table = _rowmajor(table, "table")
    ids = _i32(ids.reshape(-1), "ids")
    out = torch.empty((ids.numel(), table.size(1)), dtype=torch.float32, device=table.device)
    call("okge_gather_rows", ptr(table), _ld(table), ptr(ids), ids.numel(), table.size(1), ptr(out), out.size(1))
    return out


def scatter_add_rows(grad: torch.Tensor, ids: torch.Tensor, grad_table: torch.Tensor, skip_id: int = -1) -> None:
    grad = _rowmajor(grad, "grad")
    grad_table = _rowmajor(grad_table, "grad_table")
    ids = _i32(ids.reshape(-1), "ids")
    call("okge_scatter_add_rows", ptr(grad), _ld(grad), ptr(ids), ids.numel(), grad.size(1), int(skip_id),
         ptr(grad_table), _ld(grad_table))


def gather_pool_fwd(tok_table: torch.Tensor, id_rows: torch.Tensor, ids: Optional[torch.Tensor], mode: str,
                    id_start: int = 0, n: Optional[int] = None) -> torch.Tensor:
    tok_table = _rowmajor(tok_table, "tok_table")
    id_rows = _i32(id_rows, "id_rows")
    if ids is not None:
        ids = _i32(ids.reshape(-1), "ids")
        n = ids.numel()
    elif n is None:
        n = id_rows.size(0) - id_start
    out = torch.empty((n, tok_table.size(1)), dtype=torch.float32, device=tok_table.device)
    call("okge_gather_pool_fwd", ptr(tok_table), _ld(tok_table), ptr(id_rows), id_rows.size(1), ptr(ids),
         int(id_start), n, tok_table.size(1), POOL_MODES[mode], ptr(out), out.size(1))
    return out


def gather_pool_bwd(grad_out: torch.Tensor, tok_table: torch.Tensor, id_rows: torch.Tensor,
                    ids: Optional[torch.Tensor], mode: str, grad_tok_table: torch.Tensor, id_start: int = 0) -> None:
    grad_out = _rowmajor(grad_out, "grad_out")
    tok_table = _rowmajor(tok_table, "tok_table")
    id_rows = _i32(id_rows, "id_rows")
    if ids is not None:
        ids = _i32(ids.reshape(-1), "ids")
    if not grad_tok_table.is_contiguous() or grad_tok_table.shape != tok_table.shape:
        raise ValueError("grad_tok_table must be contiguous with the shape of tok_table")
    call("okge_gather_pool_bwd", ptr(grad_out), _ld(grad_out), ptr(tok_table), _ld(tok_table), ptr(id_rows),
         id_rows.size(1), ptr(ids), int(id_start), grad_out.size(0), grad_out.size(1), POOL_MODES[mode],
         ptr(_f32(grad_tok_table, "grad_tok_table")))


def gather_pool_bwd_slots(grad_out: torch.Tensor, tok_table: torch.Tensor, id_rows: torch.Tensor, ids: Optional[torch.Tensor],
                          mode: str, slot_map: torch.Tensor, slot_grad: torch.Tensor, id_start: int = 0) -> None:
    """``gather_pool_bwd`` into a compact gradient table: token t accumulates into ``slot_grad[slot_map[t]]``."""
    grad_out = _rowmajor(grad_out, "grad_out")
    tok_table = _rowmajor(tok_table, "tok_table")
    id_rows = _i32(id_rows, "id_rows")
    if ids is not None:
        ids = _i32(ids.reshape(-1), "ids")
    if not slot_grad.is_contiguous() or slot_grad.size(1) != tok_table.size(1) or _ld(tok_table) != tok_table.size(1):
        raise ValueError("slot_grad must be contiguous [slots, D] and tok_table contiguous")
    call("okge_gather_pool_bwd_slots", ptr(grad_out), _ld(grad_out), ptr(tok_table), _ld(tok_table), ptr(id_rows),
         id_rows.size(1), ptr(ids), int(id_start), grad_out.size(0), grad_out.size(1), POOL_MODES[mode],
         ptr(_i32(slot_map, "slot_map")), ptr(_f32(slot_grad, "slot_grad")))


def adagrad_slot_table(param: torch.Tensor, state_sum: torch.Tensor, slot_map: torch.Tensor, slot_grad: torch.Tensor,
                       clr: float, eps: float, weight_decay: float) -> None:
    """Dense Adagrad step over the whole table with the gradient taken from a compact slot table (zero where no slot)."""
    if not (param.is_contiguous() and state_sum.is_contiguous() and slot_grad.is_contiguous()):
        raise ValueError("param / state_sum / slot_grad must be contiguous")
    call("okge_adagrad_slot_table", ptr(_f32(param, "param")), ptr(_f32(state_sum, "state_sum")), param.size(0), param.size(1),
         ptr(_i32(slot_map, "slot_map")), ptr(_f32(slot_grad, "slot_grad")), float(clr), float(eps), float(weight_decay))


def dropout(x: torch.Tensor, p: float, seed: int, offset: int = 0, step_dev: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``step_dev``: int64 device scalar added (<< 44) to the stream position, for launches replayed from a CUDA graph.
    ``out``: write there (``out=x`` masks in place: the kernel is element-wise)."""
    x = _f32(x, "x")
    if not x.is_contiguous():
        if out is x:
            raise ValueError("in-place dropout needs a contiguous tensor")
        x = x.contiguous()
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("out must be a contiguous fp32 tensor of the input's shape")
    if step_dev is not None and p > 0:
        call("okge_dropout_step", ptr(x), x.numel(), float(p), int(seed) & (2**64 - 1), int(offset), ptr(step_dev), ptr(out))
    else:
        call("okge_dropout", ptr(x), x.numel(), float(p), int(seed) & (2**64 - 1), int(offset), ptr(out))
    return out


# ---------------------------------------------------------------------------------------------
# batch normalisation of encoded rows (csrc/norm_ops.cu)
# ---------------------------------------------------------------------------------------------

def _bn_workspace(n_rows: int, D: int, n_seg: int, device) -> torch.Tensor:
    nbytes = int(_capi.load().okge_bn_workspace_bytes(int(n_rows), int(D), int(n_seg)))
    return torch.empty(max(nbytes, 8) // 8 + 1, dtype=torch.float64, device=device)


def _drop_args(dropout):
    if dropout is None:
        return 0.0, 0, 0, None
    p, seed, offset, step_dev = dropout
    return float(p), int(seed) & (2**64 - 1), int(offset), ptr(step_dev)


def bn_train_fwd(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                 running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                 num_batches_tracked: Optional[torch.Tensor], momentum: float, eps: float,
                 seg: Optional[torch.Tensor] = None, n_seg: int = 1, zero_tail: bool = False, dropout=None):
    """Training-mode BatchNorm1d over the rows of ``x`` [n, D], per row segment (``seg``: int32 device tensor of n_seg
    [begin, end) pairs, None = all rows). Updates the running statistics in place; returns (y, save_mean, save_invstd).
    ``zero_tail``: rows outside the segments (padding of a fixed-capacity operand) come out as zeros. ``dropout``:
    (p, seed, offset, step_dev | None) fuses the inverted dropout that follows the normalisation into the same pass."""
    x = _rowmajor(_f32(x, "x"), "x")
    n, D = x.shape
    y = (torch.zeros if zero_tail else torch.empty)((n, D), dtype=torch.float32, device=x.device)
    save_mean = torch.empty((n_seg, D), dtype=torch.float32, device=x.device)
    save_invstd = torch.empty((n_seg, D), dtype=torch.float32, device=x.device)
    ws = _bn_workspace(n, D, n_seg, x.device)
    call("okge_bn_train_fwd", ptr(x), _ld(x), ptr(seg), int(n_seg), n, D, ptr(gamma), ptr(beta), ptr(running_mean),
         ptr(running_var), ptr(num_batches_tracked), float(momentum), float(eps), ptr(y), D, ptr(save_mean),
         ptr(save_invstd), *_drop_args(dropout), ptr(ws))
    return y, save_mean, save_invstd


def bn_train_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: Optional[torch.Tensor], save_mean: torch.Tensor,
                 save_invstd: torch.Tensor, seg: Optional[torch.Tensor] = None, n_seg: int = 1, need_dx: bool = True,
                 zero_tail: bool = False, dropout=None):
    """(dx, dgamma, dbeta) of ``bn_train_fwd``."""
    dy = _rowmajor(_f32(dy, "dy"), "dy")
    x = _rowmajor(_f32(x, "x"), "x")
    n, D = x.shape
    dx = (torch.zeros if zero_tail else torch.empty)((n, D), dtype=torch.float32, device=x.device) if need_dx else None
    dgamma = torch.empty(D, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(D, dtype=torch.float32, device=x.device)
    ws = _bn_workspace(n, D, n_seg, x.device)
    call("okge_bn_train_bwd", ptr(dy), _ld(dy), ptr(x), _ld(x), ptr(seg), int(n_seg), n, D, ptr(gamma), ptr(save_mean),
         ptr(save_invstd), ptr(dx), D, ptr(dgamma), ptr(dbeta), *_drop_args(dropout), ptr(ws))
    return dx, dgamma, dbeta


def bn_col_sums(a: torch.Tensor, x: Optional[torch.Tensor] = None, mean: Optional[torch.Tensor] = None,
                invstd: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp64 column sums [2, D] of ``a``: (sum a, sum a^2), or with ``x`` / ``mean`` / ``invstd`` given (sum a, sum a * xhat).
    Phase 1 of a batch norm whose rows are partitioned over ranks: all-reduce the result, then ``bn_normalize(_bwd)``."""
    a = _rowmajor(_f32(a, "a"), "a")
    n, D = a.shape
    sums = torch.zeros((2, D), dtype=torch.float64, device=a.device)
    if n == 0:
        return sums
    if x is not None:
        x = _rowmajor(_f32(x, "x"), "x")
    ws = _bn_workspace(n, D, 1, a.device)
    call("okge_bn_col_sums", ptr(a), _ld(a), ptr(x), _ld(x) if x is not None else 0, ptr(mean), ptr(invstd), n, D, ptr(sums),
         ptr(ws))
    return sums


def bn_normalize(x: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, gamma: Optional[torch.Tensor],
                 beta: Optional[torch.Tensor]) -> torch.Tensor:
    x = _rowmajor(_f32(x, "x"), "x")
    n, D = x.shape
    y = torch.empty((n, D), dtype=torch.float32, device=x.device)
    call("okge_bn_normalize", ptr(x), _ld(x), n, D, ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(y), D)
    return y


def bn_normalize_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, coef: torch.Tensor,
                     gamma: Optional[torch.Tensor]) -> torch.Tensor:
    """dx = gamma * invstd * (dy - coef[0] - xhat * coef[1]); ``coef`` [2, D] fp32 = (sum dy, sum dy * xhat) / n (global)."""
    dy = _rowmajor(_f32(dy, "dy"), "dy")
    x = _rowmajor(_f32(x, "x"), "x")
    n, D = x.shape
    dx = torch.empty((n, D), dtype=torch.float32, device=x.device)
    call("okge_bn_normalize_bwd", ptr(dy), _ld(dy), ptr(x), _ld(x), n, D, ptr(mean), ptr(invstd), ptr(coef), ptr(gamma),
         ptr(dx), D)
    return dx


def bn_eval_fwd(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], running_mean: torch.Tensor,
                running_var: torch.Tensor, eps: float) -> torch.Tensor:
    x = _rowmajor(_f32(x, "x"), "x")
    n, D = x.shape
    y = torch.empty((n, D), dtype=torch.float32, device=x.device)
    call("okge_bn_eval_fwd", ptr(x), _ld(x), n, D, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), float(eps),
         ptr(y), D)
    return y


# ---------------------------------------------------------------------------------------------
# LSTM token encoder: point-wise cell (csrc/lstm_ops.cu); the gate products go through gemm_nt
# ---------------------------------------------------------------------------------------------

def lstm_cell_fwd(gx: torch.Tensor, gh: Optional[torch.Tensor], b_ih: torch.Tensor, b_hh: torch.Tensor,
                  c_prev: Optional[torch.Tensor], t: int, last_state: torch.Tensor, act: Optional[torch.Tensor],
                  c: torch.Tensor, h: torch.Tensor, out: torch.Tensor) -> None:
    """One time step: gates from gx (+ gh) + biases -> act [n, 4D], c, h; rows with last_state == t copy h to out."""
    n, D = c.shape
    call("okge_lstm_cell_fwd", ptr(gx), _ld(gx), ptr(gh), _ld(gh) if gh is not None else 0, ptr(b_ih), ptr(b_hh),
         ptr(c_prev), n, D, int(t), ptr(last_state), ptr(act), ptr(c), ptr(h), ptr(out))


def lstm_cell_bwd(act: torch.Tensor, c_prev: Optional[torch.Tensor], c: torch.Tensor, grad_out: torch.Tensor,
                  last_state: torch.Tensor, t: int, dh: Optional[torch.Tensor], dc: torch.Tensor, dgates: torch.Tensor) -> None:
    """Backward of one time step: dgates [n, 4D] (pre-activation gradients), dc updated in place to d loss / d c_{t-1}."""
    n, D = c.shape
    call("okge_lstm_cell_bwd", ptr(act), ptr(c_prev), ptr(c), ptr(grad_out), ptr(last_state), int(t), ptr(dh), ptr(dc), n, D,
         ptr(dgates))


# ---------------------------------------------------------------------------------------------
# folding
# ---------------------------------------------------------------------------------------------

def fold_query(kind: int, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    a = _f32(a, "a").contiguous()
    b = _f32(b, "b").contiguous()
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError(f"fold operands must be equal-shape 2-D, got {tuple(a.shape)} and {tuple(b.shape)}")
    q = torch.empty_like(a) if out is None else out
    call("okge_fold_query", kind, ptr(a), ptr(b), a.size(0), a.size(1), ptr(q))
    return q


def fold_query_bwd(kind: int, a: torch.Tensor, b: torch.Tensor, grad_q: torch.Tensor,
                   out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    a = _f32(a, "a").contiguous()
    b = _f32(b, "b").contiguous()
    grad_q = _f32(grad_q, "grad_q").contiguous()
    ga, gb = (torch.empty_like(a), torch.empty_like(b)) if out is None else out
    call("okge_fold_query_bwd", kind, ptr(a), ptr(b), ptr(grad_q), a.size(0), a.size(1), ptr(ga), ptr(gb))
    return ga, gb


def fold_query_rows(kinds: torch.Tensor, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """ComplEx folds with one kind per row (int32 device tensor of FOLD_COMPLEX_SP / FOLD_COMPLEX_PO)."""
    a = _f32(a, "a").contiguous()
    b = _f32(b, "b").contiguous()
    q = torch.empty_like(a)
    call("okge_fold_query_rows", ptr(_i32(kinds, "kinds")), ptr(a), ptr(b), a.size(0), a.size(1), ptr(q))
    return q


def fold_query_rows_bwd(kinds: torch.Tensor, a: torch.Tensor, b: torch.Tensor, grad_q: torch.Tensor):
    a = _f32(a, "a").contiguous()
    b = _f32(b, "b").contiguous()
    grad_q = _f32(grad_q, "grad_q").contiguous()
    ga, gb = torch.empty_like(a), torch.empty_like(b)
    call("okge_fold_query_rows_bwd", ptr(_i32(kinds, "kinds")), ptr(a), ptr(b), ptr(grad_q), a.size(0), a.size(1), ptr(ga),
         ptr(gb))
    return ga, gb


# ---------------------------------------------------------------------------------------------
# fp16 operands (csrc/f16_ops.cu)
# ---------------------------------------------------------------------------------------------

_absmax_ws = {}


def _absmax_workspace(device) -> torch.Tensor:
    """Scratch of the absmax partials, one per (device, host thread). The absmax launch and the quantize launch that reads
    it are ordered on their stream; two host threads queueing on one device must not share the buffer (their launch pairs
    may interleave). A captured CUDA graph keeps reading the address it was captured with."""
    import threading
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), threading.get_ident())
    ws = _absmax_ws.get(key)
    if ws is None:
        ws = torch.zeros(256, dtype=torch.float32, device=device)
        _absmax_ws[key] = ws
    return ws


def quantize(x: torch.Tensor, split: bool = False, out: Optional[F16Operand] = None,
             fixed_scale: Optional[float] = None) -> F16Operand:
    """fp16 operand of the fp32 matrix ``x`` [rows, K] (okge_f16_absmax + okge_f16_quantize): dynamic power-of-two
    scale that puts the largest element into [128, 256) (``fixed_scale``: use that scale instead), optional lo plane.
    ``out``: write into an existing operand of the same shape (its inv_scale tensor is updated in place)."""
    x = _rowmajor(_f32(x, "x"), "x")
    rows, k = x.shape
    if out is None:
        # (the inverse scale is written by the quantize launch; only an empty operand needs the 1 filled in here)
        inv = torch.empty(1, dtype=torch.float32, device=x.device) if rows > 0 else torch.ones(1, dtype=torch.float32, device=x.device)
        out = F16Operand(torch.empty((2 if split else 1, rows, pad8(k)), dtype=torch.float16, device=x.device), rows, k, inv)
    elif out.shape != (rows, k) or (split and out.lo is None):
        raise ValueError("quantize(out=...) needs an operand of the same shape (with a lo plane for split=True)")
    if rows == 0:
        return out
    ws = None
    if fixed_scale is None:
        ws = _absmax_workspace(x.device)
        call("okge_f16_absmax", ptr(x), _ld(x), rows, k, ptr(ws))
    call("okge_f16_quantize", ptr(x), _ld(x), rows, k, ptr(ws), float(fixed_scale or 0.0), ptr(out.hi),
         ptr(out.lo) if split else None, out.ld, ptr(out.inv_scale))
    return out


def as_f16(x, split: bool = False) -> F16Operand:
    """``x`` as an fp16 operand: passed through if it already is one (with a lo plane when ``split`` asks for it)."""
    if isinstance(x, F16Operand):
        if split and x.lo is None:
            raise ValueError("split-precision contraction needs an operand quantized with split=True")
        return x if split else x.without_lo()
    return quantize(x, split=split)


def gather_rows_f16(op: F16Operand, ids: torch.Tensor) -> F16Operand:
    """Rows ``ids`` of an fp16 operand (all planes), same scale: the label / filter columns of the fused ranking."""
    ids = _i32(ids.reshape(-1), "ids")
    planes, _, ld = op.data.shape
    if op.data.stride(1) != ld or ld % 8 != 0:
        raise ValueError("gather_rows_f16 needs contiguous 16-byte rows")
    out = torch.empty((planes, ids.numel(), ld), dtype=torch.float16, device=op.device)
    for pl in range(planes):      # 16-byte words moved by the fp32 row gather
        src, dst = op.data[pl].view(torch.float32), out[pl].view(torch.float32)
        call("okge_gather_rows", ptr(src), src.stride(0), ptr(ids), ids.numel(), src.size(1), ptr(dst), dst.size(1))
    return F16Operand(out, ids.numel(), op.k, op.inv_scale)


# ---------------------------------------------------------------------------------------------
# tensor-core contractions
# ---------------------------------------------------------------------------------------------

def _operand(t: torch.Tensor, name: str) -> torch.Tensor:
    """K-major fp32 operand for TMA: unit inner stride, 16-byte aligned base, row pitch multiple of 16 B."""
    t = _rowmajor(t, name)
    if t.data_ptr() % 16 != 0 or _ld(t) % 4 != 0:
        k = t.size(1)
        buf = torch.zeros((t.size(0), pad4(k)), dtype=torch.float32, device=t.device)
        buf[:, :k] = t
        t = buf[:, :k]
    return t


def sm_count(device=None) -> int:
    return torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count


def pick_splits(M: int, N: int, K: int, k_chunk: int = 64, device=None) -> int:
    """Split K so a skinny-output contraction (dQ = dS E, K = #entities) still fills every SM."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    k_chunks = (K + k_chunk - 1) // k_chunk
    sms = sm_count(device)
    if tiles >= sms or k_chunks < 64:
        return 1
    return max(1, min(sms // tiles, k_chunks // 16))


def _is_f16(x) -> bool:
    return isinstance(x, (F16Operand, Panels, MNPanels)) or (isinstance(x, ColMajor) and isinstance(x.x, F16Operand))


def _gemm_operand_f16(x, name: str):
    """(pointer tensor, ld, layout flag, rows, K, host factor, device inverse scale) of an fp16 operand."""
    if isinstance(x, Panels):
        return x.data, 0, 1, x.rows, x.k, 1.0 / x.scale, None
    if isinstance(x, MNPanels):
        return x.data, 0, 3, x.rows, x.k, 1.0 / x.scale, None
    if isinstance(x, ColMajor):
        op = x.x if isinstance(x.x, F16Operand) else quantize(x.x)
        return op.hi, op.ld, 2, op.k, op.rows, 1.0, op.inv_scale
    op = as_f16(x)
    return op.hi, op.ld, 0, op.rows, op.k, 1.0, op.inv_scale


def _gemm_operand_f32(x, name: str):
    if isinstance(x, ColMajor):
        t = _operand(x.x, name)
        return t, _ld(t), 2, t.size(1), t.size(0)
    x = _operand(x, name)
    return x, _ld(x), 0, x.size(0), x.size(1)


def gemm_nt(a, b, alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, splits: Optional[int] = None) -> torch.Tensor:
    """out[M, N] = alpha * a[M, K] @ b[N, K]^T on the tcgen05 kernel, FP32 accumulate. fp32 tensors (optionally
    ``ColMajor``) on both sides: TF32 inputs (okge_gemm_tf32_nt). As soon as one side is an fp16 operand
    (:class:`F16Operand`, :class:`Panels`, :class:`MNPanels`, ``ColMajor(F16Operand)``) the other one is quantized too and
    the contraction runs on fp16 inputs (okge_gemm_f16_nt); scales are undone in the epilogue."""
    f16 = _is_f16(a) or _is_f16(b)
    if f16:
        at, lda, la, M, K, fa, sa = _gemm_operand_f16(a, "a")
        bt, ldb, lb, N, Kb, fb, sb = _gemm_operand_f16(b, "b")
        alpha = alpha * fa * fb
    else:
        at, lda, la, M, K = _gemm_operand_f32(a, "a")
        bt, ldb, lb, N, Kb = _gemm_operand_f32(b, "b")
    if K != Kb:
        raise ValueError(f"contraction mismatch: a is {(M, K)}, b is {(N, Kb)}")
    if out is None:
        out = torch.empty((M, pad4(N)), dtype=torch.float32, device=at.device)[:, :N]
    else:
        _rowmajor(out, "out")
    if splits is None:
        splits = pick_splits(M, N, K, 64 if f16 else 32, at.device)
    if N % 4 != 0:
        splits = 1                      # split-K partials [splits, M, N] are TMA-stored: row pitch must be 16 B aligned
    ws = None
    if splits > 1:
        ws = torch.empty((splits, M, N), dtype=torch.float32, device=at.device)
    if f16:
        call("okge_gemm_f16_nt", ptr(at), lda, la, ptr(bt), ldb, lb, M, N, K, float(alpha), ptr(alpha_dev), ptr(sa), ptr(sb),
             ptr(out), _ld(out), int(splits), ptr(ws))
    else:
        call("okge_gemm_tf32_nt", ptr(at), lda, la, ptr(bt), ldb, lb, M, N, K, float(alpha), ptr(alpha_dev), ptr(out),
             _ld(out), int(splits), ptr(ws))
    return out


def _score_operands(q, e, split: bool = False):
    q, e = as_f16(q, split), as_f16(e, split)
    if q.k != e.k:
        raise ValueError(f"width mismatch: q is {q.shape}, e is {e.shape}")
    return q, e


def _lo(op: F16Operand, split: bool):
    return ptr(op.lo) if split else None


def score_store(q, e, split: bool = False) -> torch.Tensor:
    """scores [B, N] = q e^T materialised. ``split``: three-term split-precision product (both operands carry lo planes)."""
    q, e = _score_operands(q, e, split)
    B, D = q.shape
    N = e.rows
    out = torch.empty((B, pad4(N)), dtype=torch.float32, device=q.device)[:, :N]
    call("okge_score_store", ptr(q.hi), _lo(q, split), q.ld, ptr(e.hi), _lo(e, split), e.ld, B, N, D, ptr(q.inv_scale),
         ptr(e.inv_scale), ptr(out), _ld(out))
    return out


def score_bce(q, e, pos_ptr: torch.Tensor, pos_idx: torch.Tensor, y_base: float = 0.0, y_pos: float = 1.0,
              want_dS: bool = True, n_cols_dev: Optional[torch.Tensor] = None, ds_scale: float = DS_SCALE_BCE):
    """Returns (loss_sum [1] float64 device tensor, dS [B, N] | None); the gradient is a :class:`Panels` operand (the
    layout the dQ / dE contractions read). ``q`` / ``e``: fp32 matrices (quantized here) or :class:`F16Operand`.
    ``n_cols_dev``: int32 device scalar, number of real candidates when ``e`` is padded to a fixed capacity (columns
    behind it: no loss, zero gradient)."""
    q, e = _score_operands(q, e)
    B, D = q.shape
    N = e.rows
    pos_ptr = _i32(pos_ptr, "pos_ptr")
    pos_idx = _i32(pos_idx, "pos_idx")
    loss = torch.empty(1, dtype=torch.float64, device=q.device)
    dS = Panels.empty(B, N, q.device, ds_scale) if want_dS else None
    call("okge_score_bce", ptr(q.hi), q.ld, ptr(e.hi), e.ld, B, N, D, ptr(q.inv_scale), ptr(e.inv_scale), ptr(pos_ptr),
         ptr(pos_idx), float(y_base), float(y_pos), ptr(n_cols_dev), ptr(loss), ptr(dS.data) if dS is not None else None,
         float(ds_scale))
    return loss, dS


def score_bce_rank(q, e, pos_ptr: torch.Tensor, pos_idx: torch.Tensor, y_base: float, y_pos: float,
                   thresh4: torch.Tensor, greater4: torch.Tensor, equal4: torch.Tensor, loss_out: torch.Tensor,
                   extra_rows: int = 0, split: bool = False, slots: int = 4) -> None:
    """Evaluation pass: BCE loss sum into ``loss_out`` [1] float64 AND the count-greater / count-equal of up to 4 ranked
    answers per query row (``thresh4`` [B, 4], +inf = unused; counts added to ``greater4`` / ``equal4`` [B, 4] int32).
    The last ``extra_rows`` rows of ``q`` (and of the three [.., 4] arrays) only rank (no loss term). ``slots``: how many
    of the 4 slots per row the kernel looks at (1, 2 or 4; each costs four instructions per score)."""
    q, e = _score_operands(q, e, split)
    B, D = q.rows - int(extra_rows), q.k
    call("okge_score_bce_rank", ptr(q.hi), _lo(q, split), q.ld, ptr(e.hi), _lo(e, split), e.ld, B, int(extra_rows), e.rows, D,
         ptr(q.inv_scale), ptr(e.inv_scale), ptr(_i32(pos_ptr, "pos_ptr")), ptr(_i32(pos_idx, "pos_idx")), float(y_base),
         float(y_pos), ptr(thresh4), ptr(greater4), ptr(equal4), int(slots), ptr(loss_out))


def score_lse(q, e, pos_ptr: torch.Tensor, pos_idx: torch.Tensor):
    """Returns (row_lse [B], pos_score [nnz])."""
    q, e = _score_operands(q, e)
    B, D = q.shape
    N = e.rows
    pos_ptr = _i32(pos_ptr, "pos_ptr")
    pos_idx = _i32(pos_idx, "pos_idx")
    row_lse = torch.empty(B, dtype=torch.float32, device=q.device)
    pos_score = torch.zeros(max(pos_idx.numel(), 1), dtype=torch.float32, device=q.device)
    ws = torch.empty(_capi.load().okge_score_lse_ws_floats(B, N), dtype=torch.float32, device=q.device)
    call("okge_score_lse", ptr(q.hi), q.ld, ptr(e.hi), e.ld, B, N, D, ptr(q.inv_scale), ptr(e.inv_scale), ptr(pos_ptr),
         ptr(pos_idx), ptr(row_lse), ptr(pos_score), ptr(ws))
    return row_lse, pos_score[: pos_idx.numel()]


def score_softmax_grad(q, e, pos_ptr: torch.Tensor, pos_idx: torch.Tensor, row_lse: torch.Tensor,
                       row_weight: torch.Tensor, ds_scale: float = DS_SCALE_KL) -> Panels:
    q, e = _score_operands(q, e)
    B, D = q.shape
    N = e.rows
    pos_ptr = _i32(pos_ptr, "pos_ptr")
    pos_idx = _i32(pos_idx, "pos_idx")
    dS = Panels.empty(B, N, q.device, ds_scale)
    call("okge_score_softmax_grad", ptr(q.hi), q.ld, ptr(e.hi), e.ld, B, N, D, ptr(q.inv_scale), ptr(e.inv_scale),
         ptr(pos_ptr), ptr(pos_idx), ptr(_f32(row_lse, "row_lse").contiguous()),
         ptr(_f32(row_weight, "row_weight").contiguous()), ptr(dS.data), float(ds_scale))
    return dS


# ---------------------------------------------------------------------------------------------
# ranking
# ---------------------------------------------------------------------------------------------

def rank_count(scores: torch.Tensor, ans_row: torch.Tensor, alt_ptr: torch.Tensor, alt_idx: torch.Tensor,
               filt_ptr: torch.Tensor, filt_idx: torch.Tensor):
    """(true_score [Q] f32, greater [Q] i32, equal [Q] i32) over a materialised score matrix."""
    scores = _rowmajor(scores, "scores")
    B, N = scores.shape
    ans_row, alt_ptr, alt_idx = _i32(ans_row, "ans_row"), _i32(alt_ptr, "alt_ptr"), _i32(alt_idx, "alt_idx")
    filt_ptr, filt_idx = _i32(filt_ptr, "filt_ptr"), _i32(filt_idx, "filt_idx")
    Q = ans_row.numel()
    true = torch.empty(Q, dtype=torch.float32, device=scores.device)
    greater = torch.empty(Q, dtype=torch.int32, device=scores.device)
    equal = torch.empty(Q, dtype=torch.int32, device=scores.device)
    call("okge_rank_count", ptr(scores), _ld(scores), B, N, ptr(ans_row), ptr(alt_ptr), ptr(alt_idx), Q,
         ptr(filt_ptr), ptr(filt_idx), ptr(true), ptr(greater), ptr(equal))
    return true, greater, equal


def score_rank(q, e, thresh: torch.Tensor, greater: torch.Tensor, equal: torch.Tensor, split: bool = False) -> None:
    q, e = _score_operands(q, e, split)
    call("okge_score_rank", ptr(q.hi), _lo(q, split), q.ld, ptr(e.hi), _lo(e, split), e.ld, q.rows, e.rows, q.k,
         ptr(q.inv_scale), ptr(e.inv_scale), ptr(_f32(thresh, "thresh").contiguous()), ptr(greater), ptr(equal))


def rank_true_score(sel_scores: torch.Tensor, ans_row: torch.Tensor, alt_ptr: torch.Tensor, alt_pos: torch.Tensor,
                    true_score: torch.Tensor) -> None:
    sel_scores = _rowmajor(sel_scores, "sel_scores")
    call("okge_rank_true_score", ptr(sel_scores), _ld(sel_scores), ptr(_i32(ans_row, "ans_row")),
         ptr(_i32(alt_ptr, "alt_ptr")), ptr(_i32(alt_pos, "alt_pos")), ans_row.numel(), ptr(true_score))


def rank_filter_correct(sel_scores: torch.Tensor, ans_row: torch.Tensor, filt_ptr: torch.Tensor, filt_pos: torch.Tensor,
                        thresh: torch.Tensor, greater: torch.Tensor, equal: torch.Tensor, add_mask_terms: bool = True) -> None:
    sel_scores = _rowmajor(sel_scores, "sel_scores")
    call("okge_rank_filter_correct", ptr(sel_scores), _ld(sel_scores), ptr(_i32(ans_row, "ans_row")), ans_row.numel(),
         ptr(_i32(filt_ptr, "filt_ptr")), ptr(_i32(filt_pos, "filt_pos")), ptr(thresh), int(bool(add_mask_terms)),
         ptr(greater), ptr(equal))


# ---------------------------------------------------------------------------------------------
# optimizers
# ---------------------------------------------------------------------------------------------

def _flat(t: torch.Tensor, name: str) -> torch.Tensor:
    _f32(t, name)
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous (updated in place)")
    return t


def adagrad_dense(param, grad, state_sum, clr: float, eps: float, weight_decay: float) -> None:
    call("okge_adagrad_dense", ptr(_flat(param, "param")), ptr(_flat(grad, "grad")), ptr(_flat(state_sum, "state_sum")),
         param.numel(), float(clr), float(eps), float(weight_decay))


def adagrad_rows(param, state_sum, grad_rows, row_ids, clr: float, eps: float, weight_decay: float = 0.0,
                 slot_map: Optional[torch.Tensor] = None) -> None:
    """Row-wise Adagrad on rows ``row_ids`` (unique, or with repeats when ``slot_map`` names the owning position)."""
    grad_rows = _rowmajor(grad_rows, "grad_rows")
    call("okge_adagrad_rows", ptr(_flat(param, "param")), ptr(_flat(state_sum, "state_sum")), param.size(1),
         ptr(grad_rows), _ld(grad_rows), ptr(_i32(row_ids, "row_ids")), ptr(slot_map), row_ids.numel(), param.size(1), float(clr),
         float(eps), float(weight_decay))


def gemm_adagrad(a, b, param: torch.Tensor, state_sum: torch.Tensor, clr: float, eps: float, weight_decay: float,
                 alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None, extra_map: Optional[torch.Tensor] = None,
                 extra: Optional[torch.Tensor] = None, shadow: Optional[F16Operand] = None, dropout=None) -> None:
    """param, state_sum <- Adagrad(param, alpha * a @ b^T + extra[extra_map], state_sum) in one pass of the tensor-core
    kernel (the gradient never reaches memory). ``param`` / ``state_sum``: [M, N] row-major views updated in place;
    ``a`` / ``b``: MN-major fp16 operands (``dS.T``, ``ColMajor(q16)``). ``shadow``: fp16 operand of the same rows that
    receives fp16(param_new * its scale) -- the scoring operand of the next step, kept current by the update itself.
    ``dropout`` = (p, seed, offset, step_dev): the contraction is the gradient of the dropped-out rows
    (``mask_dropout_f16`` with the same arguments); the mask is applied to the gradient tile."""
    at, lda, la, M, K, fa, sa = _gemm_operand_f16(a, "a")
    bt, ldb, lb, N, Kb, fb, sb = _gemm_operand_f16(b, "b")
    if K != Kb:
        raise ValueError(f"contraction mismatch: a is {(M, K)}, b is {(N, Kb)}")
    for t, name in ((param, "param"), (state_sum, "state_sum")):
        _f32(t, name)
        if tuple(t.shape) != (M, N) or t.stride(1) != 1 or t.stride(0) != param.stride(0):
            raise ValueError(f"{name} must be a [{M}, {N}] row-major view (shared row pitch), got {tuple(t.shape)}")
    if extra_map is not None:
        extra_map = _i32(extra_map, "extra_map")
        extra = _rowmajor(extra, "extra")
    sh_ptr, sh_ld, sh_inv = None, 0, None
    if shadow is not None:
        if shadow.shape != (M, N) or shadow.inv_scale is None:
            raise ValueError(f"shadow must be an fp16 operand of shape {(M, N)} with a device scale")
        sh_ptr, sh_ld, sh_inv = ptr(shadow.hi), shadow.ld, ptr(shadow.inv_scale)
    args = (ptr(at), lda, la, ptr(bt), ldb, lb, M, N, K, float(alpha * fa * fb), ptr(alpha_dev), ptr(sa),
            ptr(sb), ptr(extra_map), ptr(extra), _ld(extra) if extra is not None else 0, ptr(param), ptr(state_sum),
            param.stride(0), sh_ptr, sh_ld, sh_inv, float(clr), float(eps), float(weight_decay))
    if dropout is None:
        call("okge_gemm_adagrad", *args)
    else:
        dp, seed, offset, step_dev = dropout
        call("okge_gemm_adagrad_dropout", *args, float(dp), int(seed) & (2 ** 64 - 1), int(offset), ptr(step_dev))


def mask_dropout_f16(op: F16Operand, p: float, seed: int, offset: int = 0, step_dev: Optional[torch.Tensor] = None,
                     out: Optional[F16Operand] = None) -> F16Operand:
    """The fp16 operand of inverted-dropout(x) from the fp16 operand of x: elements the mask of ``dropout(x, p, seed,
    offset, step_dev)`` drops become zero, the 1 / (1 - p) goes into the (new) inverse scale. hi plane only."""
    if op.k % 4 != 0 or op.ld % 4 != 0:
        raise ValueError("mask_dropout_f16 needs a column count and row pitch that are multiples of 4")
    if out is None:
        out = F16Operand(torch.empty((1, op.rows, op.ld), dtype=torch.float16, device=op.device), op.rows, op.k,
                         torch.empty(1, dtype=torch.float32, device=op.device))
    call("okge_f16_mask_dropout", ptr(op.hi), op.ld, op.rows, op.k, float(p), int(seed) & (2 ** 64 - 1), int(offset), ptr(step_dev),
         ptr(op.inv_scale), ptr(out.hi), out.ld, ptr(out.inv_scale))
    return out


def batch_layout(n_po_dev: torch.Tensor, rows: int, n_cols: int, kind_po: int, kind_sp: int, kinds: Optional[torch.Tensor],
                 segments: Optional[torch.Tensor], token_model: bool, count_dev: Optional[torch.Tensor] = None,
                 step_counter: Optional[torch.Tensor] = None) -> None:
    """okge_batch_layout: row kinds / batch-norm segments of a batch from the device copy of its number of po rows (and
    the dropout step counter += 1), one launch inside a captured step."""
    if n_po_dev.dtype != torch.int32 or (count_dev is not None and count_dev.dtype != torch.int32):
        raise TypeError("n_po_dev / count_dev must be int32 device tensors")
    if step_counter is not None and step_counter.dtype != torch.int64:
        raise TypeError("step_counter must be an int64 device tensor")
    call("okge_batch_layout", ptr(n_po_dev), ptr(count_dev), int(rows), int(n_cols), int(kind_po), int(kind_sp),
         ptr(_i32(kinds, "kinds")) if kinds is not None else None, ptr(_i32(segments, "segments")) if segments is not None else None,
         1 if token_model else 0, ptr(step_counter))


def row_slots_build(ids: torch.Tensor, slot_map: torch.Tensor, skip_id: int = -1) -> None:
    ids = _i32(ids.reshape(-1), "ids")
    call("okge_row_slots_build", ptr(ids), ids.numel(), int(skip_id), ptr(_i32(slot_map, "slot_map")))


def row_slots_accumulate(grad: torch.Tensor, ids: torch.Tensor, slot_map: torch.Tensor, extra: torch.Tensor,
                         skip_id: int = -1) -> None:
    grad = _rowmajor(grad, "grad")
    extra = _rowmajor(extra, "extra")
    ids = _i32(ids.reshape(-1), "ids")
    call("okge_row_slots_accumulate", ptr(grad), _ld(grad), ptr(ids), ids.numel(), grad.size(1), int(skip_id),
         ptr(_i32(slot_map, "slot_map")), ptr(extra), _ld(extra))


def adagrad_slot_rows(param: torch.Tensor, state_sum: torch.Tensor, n_rows: int, slot_map: Optional[torch.Tensor],
                      extra: Optional[torch.Tensor], clr: float, eps: float, weight_decay: float) -> None:
    """Adagrad on rows [0, n_rows) of ``param`` with gradient ``extra[slot_map[r]]`` (zero where the slot is -1)."""
    call("okge_adagrad_slot_rows", ptr(_flat(param, "param")), ptr(_flat(state_sum, "state_sum")), param.size(1), int(n_rows),
         param.size(1), ptr(slot_map), ptr(extra), _ld(extra) if extra is not None else 0, float(clr), float(eps),
         float(weight_decay))


def row_slots_clear(ids: torch.Tensor, slot_map: torch.Tensor, skip_id: int = -1) -> None:
    ids = _i32(ids.reshape(-1), "ids")
    call("okge_row_slots_clear", ptr(ids), ids.numel(), int(skip_id), ptr(_i32(slot_map, "slot_map")))


def adam_dense(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step: int) -> None:
    call("okge_adam_dense", ptr(_flat(param, "param")), ptr(_flat(grad, "grad")), ptr(_flat(exp_avg, "exp_avg")),
         ptr(_flat(exp_avg_sq, "exp_avg_sq")), param.numel(), float(lr), float(beta1), float(beta2), float(eps),
         float(weight_decay), 1.0 - beta1 ** step, 1.0 - beta2 ** step)


def adam_rows(param, exp_avg, exp_avg_sq, grad_rows, row_ids, lr, beta1, beta2, eps, weight_decay, step: int,
              slot_map: Optional[torch.Tensor] = None) -> None:
    grad_rows = _rowmajor(grad_rows, "grad_rows")
    call("okge_adam_rows", ptr(_flat(param, "param")), ptr(_flat(exp_avg, "exp_avg")), ptr(_flat(exp_avg_sq, "exp_avg_sq")),
         param.size(1), ptr(grad_rows), _ld(grad_rows), ptr(_i32(row_ids, "row_ids")), ptr(slot_map), row_ids.numel(),
         param.size(1), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), 1.0 - beta1 ** step,
         1.0 - beta2 ** step)


# ---------------------------------------------------------------------------------------------
# batch-shared collate on the device
# ---------------------------------------------------------------------------------------------

COLLATE_B_PO, COLLATE_COUNT, COLLATE_NNZ, COLLATE_N_UNIQUE, COLLATE_OVERFLOW, COLLATE_NNZ_TOTAL, COLLATE_CALLS = range(7)
COLLATE_SCALARS = 8


def collate_shared(rows, lab_ptr, lab_idx, prefix, slot, n_entities: int, id_offset: int, min_size: int, cap_nnz: int,
                   cap_cols: int, n_draw: int, seed: int, ws: dict, out: dict) -> None:
    """okge_collate_shared: ``rows`` (device int64 [B]) of the prefix index -> the tensors of a batch-shared training batch,
    written into the preallocated ``out`` (ent, rel, is_po, ptr, idx, cand, scalars, count, inv_norm) using the workspace
    ``ws`` (bitmap, word_prefix, tile_sum, first_draw, e_flat, row_start); see include/okge_b200.h."""
    if rows.dtype != torch.int64 or not rows.is_cuda or not rows.is_contiguous():
        raise TypeError("rows must be a contiguous CUDA int64 tensor")
    B = rows.numel()
    n_words = (n_entities + 31) // 32
    need = {"bitmap": n_words, "word_prefix": n_words, "tile_sum": (n_words + 1023) // 1024, "first_draw": n_entities,
            "e_flat": cap_nnz + n_draw, "row_start": B}
    for k, n in need.items():
        if ws[k].numel() < n or not ws[k].is_cuda or not ws[k].is_contiguous():
            raise ValueError(f"workspace {k} needs {n} contiguous device elements")
    sizes = {"ent": B, "rel": B, "is_po": B, "ptr": B + 1, "idx": cap_nnz, "cand": cap_cols, "scalars": COLLATE_SCALARS}
    for k, n in sizes.items():
        if out[k].numel() < n or not out[k].is_cuda or not out[k].is_contiguous():
            raise ValueError(f"output {k} needs {n} contiguous device elements")
        if out[k].dtype != (torch.int64 if k == "scalars" else torch.int32):
            raise TypeError(f"output {k} has dtype {out[k].dtype}")
    call("okge_collate_shared", ptr(rows), B, ptr(lab_ptr), ptr(lab_idx), ptr(prefix), ptr(slot), int(n_entities),
         int(id_offset), int(min_size), int(cap_nnz), int(cap_cols), int(n_draw), int(seed) & (2 ** 64 - 1),
         ptr(ws["bitmap"]), ptr(ws["word_prefix"]), ptr(ws["tile_sum"]), ptr(ws["first_draw"]), ptr(ws["e_flat"]),
         ptr(ws["row_start"]), ptr(out["ent"]), ptr(out["rel"]), ptr(out["is_po"]), ptr(out["ptr"]), ptr(out["idx"]),
         ptr(out["cand"]), ptr(out["scalars"]), ptr(out.get("count")), ptr(out.get("inv_norm")))

