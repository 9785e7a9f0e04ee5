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
    "transpose", "rank_count", "score_rank", "rank_true_score", "rank_filter_correct",
    "adagrad_dense", "adagrad_rows", "adam_dense", "adam_rows", "gemm_adagrad", "row_slots_build",
    "row_slots_accumulate", "row_slots_clear", "adagrad_slot_rows", "bn_train_fwd", "bn_train_bwd", "bn_eval_fwd", "bn_col_sums", "bn_normalize", "bn_normalize_bwd", "lstm_cell_fwd", "lstm_cell_bwd", "pad4", "Panels", "MNPanels", "ColMajor",
    "transposed_operand", "TF32_RAW_OPERAND_SCALE",
    "FOLD_COMPLEX_SP", "FOLD_COMPLEX_PO", "FOLD_DISTMULT",
]

SM_COUNT_B200 = 148
# tcgen05 truncates raw fp32 operands to TF32; 1 / (1 - 2^-11 / (2 ln 2)) centres the error (okge_common.cuh)
TF32_RAW_OPERAND_SCALE = 1.0 / (1.0 - 0.00035221)


def pad4(n: int) -> int:
    """Leading dimensions feeding the tensor-core path must be multiples of 4 floats (16 B)."""
    return (int(n) + 3) // 4 * 4


class Panels:
    """A logical [rows, K] fp32 GEMM operand stored as K-panels ``[ceil(K/32), rows, 32]`` (OKGE_K_PANELS in
    okge_b200.h): every TMA box of the tensor-core kernel is one contiguous block of memory. Produced by the
    loss epilogues (dS, dST) and by ``transposed_operand``; consumed by ``gemm_nt``."""

    def __init__(self, data: torch.Tensor, rows: int, k: int):
        self.data, self.rows, self.k = data, int(rows), int(k)
        self.shape = (self.rows, self.k)
        self.device = data.device

    @staticmethod
    def empty(rows: int, k: int, device) -> "Panels":
        return Panels(torch.empty(((k + 31) // 32, rows, 32), dtype=torch.float32, device=device), rows, k)

    def size(self, dim: int) -> int:
        return self.shape[dim]

    def dense(self) -> torch.Tensor:
        """The logical [rows, K] matrix (tests / debugging)."""
        return self.data.permute(1, 0, 2).reshape(self.rows, -1)[:, : self.k].contiguous()

    @property
    def T(self) -> "MNPanels":
        """The transposed operand [K, rows] over the SAME storage (OKGE_MN_PANELS): no data movement."""
        return MNPanels(self.data, self.k, self.rows)


class MNPanels:
    """A logical [rows, K] operand stored as ``[ceil(rows/32), K, 32]`` (OKGE_MN_PANELS): the K-panel storage of the
    transposed matrix, read MN-major by the tensor core. ``dS.T`` is how dE = dS^T Q consumes the loss gradient."""

    def __init__(self, data: torch.Tensor, rows: int, k: int):
        self.data, self.rows, self.k = data, int(rows), int(k)
        self.shape = (self.rows, self.k)
        self.device = data.device

    def size(self, dim: int) -> int:
        return self.shape[dim]


class ColMajor:
    """``ColMajor(x)`` presents a row-major ``x[K, rows]`` as the logical operand ``x^T [rows, K]`` (OKGE_COL_MAJOR)
    without a transpose pass: E[N, D] enters dQ = dS E and Q[B, D] enters dE = dS^T Q this way."""

    def __init__(self, x: torch.Tensor):
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


def dropout(x: torch.Tensor, p: float, seed: int, offset: int = 0, step_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``step_dev``: int64 device scalar added (<< 44) to the stream position, for launches replayed from a CUDA graph."""
    x = _f32(x, "x").contiguous()
    out = torch.empty_like(x)
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
# tensor-core contractions
# ---------------------------------------------------------------------------------------------

def _operand(t: torch.Tensor, name: str) -> torch.Tensor:
    """K-major operand for TMA: unit inner stride, 16-byte aligned base, row pitch multiple of 16 B."""
    t = _rowmajor(t, name)
    if t.data_ptr() % 16 != 0 or _ld(t) % 4 != 0:
        k = t.size(1)
        buf = torch.zeros((t.size(0), pad4(k)), dtype=torch.float32, device=t.device)
        buf[:, :k] = t
        t = buf[:, :k]
    return t


def pick_splits(M: int, N: int, K: int) -> int:
    """Split K so a skinny-output contraction (dQ = dS E, K = #entities) still fills 148 SMs."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    k_chunks = (K + 31) // 32
    if tiles >= SM_COUNT_B200 or k_chunks < 64:
        return 1
    return max(1, min(SM_COUNT_B200 // tiles, k_chunks // 16))


def _gemm_operand(x, name: str):
    """(pointer tensor, ld, layout flag, rows, K) of a row-major tensor or a Panels operand."""
    if isinstance(x, Panels):
        return x.data, 0, 1, x.rows, x.k
    if isinstance(x, MNPanels):
        return x.data, 0, 3, x.rows, x.k
    if isinstance(x, ColMajor):
        t = _operand(x.x, name)
        return t, _ld(t), 2, t.size(1), t.size(0)
    x = _operand(x, name)
    return x, _ld(x), 0, x.size(0), x.size(1)


def gemm_nt(a, b, alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, splits: Optional[int] = None) -> torch.Tensor:
    """out[M, N] = alpha * a[M, K] @ b[N, K]^T on the tcgen05 kernel (TF32 in, FP32 accumulate); ``a`` / ``b``
    are row-major tensors or :class:`Panels`."""
    at, lda, la, M, K = _gemm_operand(a, "a")
    bt, ldb, lb, N, Kb = _gemm_operand(b, "b")
    if K != Kb:
        raise ValueError(f"contraction mismatch: a is {(M, K)}, b is {(N, Kb)}")
    if out is None:
        out = torch.empty((M, pad4(N)), dtype=torch.float32, device=at.device)[:, :N]
    else:
        _rowmajor(out, "out")
    if splits is None:
        splits = pick_splits(M, N, K)
    if N % 4 != 0:
        splits = 1                      # split-K partials [splits, M, N] are TMA-stored: row pitch must be 16 B aligned
    ws = None
    if splits > 1:
        ws = torch.empty((splits, M, N), dtype=torch.float32, device=at.device)
    call("okge_gemm_tf32_nt", ptr(at), lda, la, ptr(bt), ldb, lb, M, N, K, float(alpha), ptr(alpha_dev), ptr(out),
         _ld(out), int(splits), ptr(ws))
    return out


def score_store(q: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    q = _operand(q, "q")
    e = _operand(e, "e")
    B, D = q.shape
    N = e.size(0)
    out = torch.empty((B, pad4(N)), dtype=torch.float32, device=q.device)[:, :N]
    call("okge_score_store", ptr(q), _ld(q), ptr(e), _ld(e), B, N, D, ptr(out), _ld(out))
    return out


def score_bce(q: torch.Tensor, e: torch.Tensor, pos_ptr: torch.Tensor, pos_idx: torch.Tensor, y_base: float = 0.0,
              y_pos: float = 1.0, want_dS: bool = True, want_dST: bool = True, n_cols_dev: Optional[torch.Tensor] = None):
    """Returns (loss_sum [1] float64 device tensor, dS [B, N] | None, dST [N, B] | None); the gradients are
    :class:`Panels` (the layout the dQ / dE contractions read). ``n_cols_dev``: int32 device scalar, number of real
    candidates when ``e`` is padded to a fixed capacity (columns behind it: no loss, zero gradient)."""
    q = _operand(q, "q")
    e = _operand(e, "e")
    B, D = q.shape
    N = e.size(0)
    pos_ptr = _i32(pos_ptr, "pos_ptr")
    pos_idx = _i32(pos_idx, "pos_idx")
    loss = torch.empty(1, dtype=torch.float64, device=q.device)
    dS = Panels.empty(B, N, q.device) if want_dS else None
    dST = Panels.empty(N, B, q.device) if want_dST else None
    call("okge_score_bce", ptr(q), _ld(q), ptr(e), _ld(e), B, N, D, ptr(pos_ptr), ptr(pos_idx), float(y_base),
         float(y_pos), ptr(n_cols_dev), ptr(loss), ptr(dS.data) if dS is not None else None,
         ptr(dST.data) if dST is not None else None)
    return loss, dS, dST


def score_bce_rank(q: torch.Tensor, e: torch.Tensor, pos_ptr: torch.Tensor, pos_idx: torch.Tensor, y_base: float, y_pos: float,
                   thresh4: torch.Tensor, greater4: torch.Tensor, equal4: torch.Tensor, loss_out: torch.Tensor,
                   extra_rows: int = 0) -> None:
    """Evaluation pass: BCE loss sum into ``loss_out`` [1] float64 AND the count-greater / count-equal of up to 4 ranked
    answers per query row (``thresh4`` [B, 4], +inf = unused; counts added to ``greater4`` / ``equal4`` [B, 4] int32).
    The last ``extra_rows`` rows of ``q`` (and of the three [.., 4] arrays) only rank (no loss term)."""
    q = _operand(q, "q")
    e = _operand(e, "e")
    B, D = q.size(0) - int(extra_rows), q.size(1)
    call("okge_score_bce_rank", ptr(q), _ld(q), ptr(e), _ld(e), B, int(extra_rows), e.size(0), D, ptr(_i32(pos_ptr, "pos_ptr")),
         ptr(_i32(pos_idx, "pos_idx")), float(y_base), float(y_pos), ptr(thresh4), ptr(greater4), ptr(equal4), ptr(loss_out))


def score_lse(q: torch.Tensor, e: torch.Tensor, pos_ptr: torch.Tensor, pos_idx: torch.Tensor):
    """Returns (row_lse [B], pos_score [nnz])."""
    q = _operand(q, "q")
    e = _operand(e, "e")
    B, D = q.shape
    N = e.size(0)
    pos_ptr = _i32(pos_ptr, "pos_ptr")
    pos_idx = _i32(pos_idx, "pos_idx")
    row_lse = torch.empty(B, dtype=torch.float32, device=q.device)
    pos_score = torch.zeros(max(pos_idx.numel(), 1), dtype=torch.float32, device=q.device)
    ws = torch.empty(_capi.load().okge_score_lse_ws_floats(B, N), dtype=torch.float32, device=q.device)
    call("okge_score_lse", ptr(q), _ld(q), ptr(e), _ld(e), B, N, D, ptr(pos_ptr), ptr(pos_idx), ptr(row_lse),
         ptr(pos_score), ptr(ws))
    return row_lse, pos_score[: pos_idx.numel()]


def score_softmax_grad(q: torch.Tensor, e: torch.Tensor, pos_ptr: torch.Tensor, pos_idx: torch.Tensor,
                       row_lse: torch.Tensor, row_weight: torch.Tensor, want_dS: bool = True, want_dST: bool = True):
    q = _operand(q, "q")
    e = _operand(e, "e")
    B, D = q.shape
    N = e.size(0)
    pos_ptr = _i32(pos_ptr, "pos_ptr")
    pos_idx = _i32(pos_idx, "pos_idx")
    dS = Panels.empty(B, N, q.device) if want_dS else None
    dST = Panels.empty(N, B, q.device) if want_dST else None
    call("okge_score_softmax_grad", ptr(q), _ld(q), ptr(e), _ld(e), B, N, D, ptr(pos_ptr), ptr(pos_idx),
         ptr(_f32(row_lse, "row_lse").contiguous()), ptr(_f32(row_weight, "row_weight").contiguous()),
         ptr(dS.data) if dS is not None else None, ptr(dST.data) if dST is not None else None)
    return dS, dST


def transpose(x: torch.Tensor, round_tf32: bool = False) -> torch.Tensor:
    """Returns x^T as a fresh K-major operand ([cols, rows], row pitch padded to 16 bytes); with
    ``round_tf32`` the values are rounded to nearest TF32 (exact under the tensor core's truncation)."""
    x = _rowmajor(x, "x")
    rows, cols = x.shape
    out = torch.empty((cols, pad4(rows)), dtype=torch.float32, device=x.device)[:, :rows]
    call("okge_transpose", ptr(x), _ld(x), rows, cols, ptr(out), _ld(out), int(bool(round_tf32)))
    return out


def transposed_operand(x: torch.Tensor, round_tf32: bool = True) -> Panels:
    """x^T as a K-panel operand: logical [cols, K = rows] — how the candidate table E[N, D] enters dQ = dS E."""
    x = _rowmajor(x, "x")
    rows, cols = x.shape
    out = Panels.empty(cols, rows, x.device)
    call("okge_transpose_to_panels", ptr(x), _ld(x), rows, cols, ptr(out.data), int(bool(round_tf32)))
    return out


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


def score_rank(q: torch.Tensor, e: torch.Tensor, thresh: torch.Tensor, greater: torch.Tensor, equal: torch.Tensor) -> None:
    q = _operand(q, "q")
    e = _operand(e, "e")
    Q, D = q.shape
    call("okge_score_rank", ptr(q), _ld(q), ptr(e), _ld(e), Q, e.size(0), D,
         ptr(_f32(thresh, "thresh").contiguous()), ptr(greater), ptr(equal))


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


def adagrad_rows(param, state_sum, grad_rows, row_ids, clr: float, eps: float, weight_decay: float = 0.0) -> None:
    grad_rows = _rowmajor(grad_rows, "grad_rows")
    call("okge_adagrad_rows", ptr(_flat(param, "param")), ptr(_flat(state_sum, "state_sum")), param.size(1),
         ptr(grad_rows), _ld(grad_rows), ptr(_i32(row_ids, "row_ids")), row_ids.numel(), param.size(1), float(clr),
         float(eps), float(weight_decay))


def gemm_adagrad(a, b, param: torch.Tensor, state_sum: torch.Tensor, clr: float, eps: float, weight_decay: float,
                 alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None, extra_map: Optional[torch.Tensor] = None,
                 extra: Optional[torch.Tensor] = None) -> None:
    """param, state_sum <- Adagrad(param, alpha * a @ b^T + extra[extra_map], state_sum) in one pass of the tensor-core
    kernel (the gradient never reaches memory). ``param`` / ``state_sum``: [M, N] row-major views updated in place."""
    at, lda, la, M, K = _gemm_operand(a, "a")
    bt, ldb, lb, N, Kb = _gemm_operand(b, "b")
    if K != Kb:
        raise ValueError(f"contraction mismatch: a is {(M, K)}, b is {(N, Kb)}")
    for t, name in ((param, "param"), (state_sum, "state_sum")):
        _f32(t, name)
        if tuple(t.shape) != (M, N) or t.stride(1) != 1 or t.stride(0) != param.stride(0):
            raise ValueError(f"{name} must be a [{M}, {N}] row-major view (shared row pitch), got {tuple(t.shape)}")
    if extra_map is not None:
        extra_map = _i32(extra_map, "extra_map")
        extra = _rowmajor(extra, "extra")
    call("okge_gemm_adagrad", ptr(at), lda, la, ptr(bt), ldb, lb, M, N, K, float(alpha), ptr(alpha_dev), ptr(extra_map),
         ptr(extra), _ld(extra) if extra is not None else 0, ptr(param), ptr(state_sum), param.stride(0), float(clr),
         float(eps), float(weight_decay))


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


def adam_rows(param, exp_avg, exp_avg_sq, grad_rows, row_ids, lr, beta1, beta2, eps, weight_decay, step: int) -> None:
    grad_rows = _rowmajor(grad_rows, "grad_rows")
    call("okge_adam_rows", ptr(_flat(param, "param")), ptr(_flat(exp_avg, "exp_avg")), ptr(_flat(exp_avg_sq, "exp_avg_sq")),
         param.size(1), ptr(grad_rows), _ld(grad_rows), ptr(_i32(row_ids, "row_ids")), row_ids.numel(), param.size(1),
         float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), 1.0 - beta1 ** step,
         1.0 - beta2 ** step)
