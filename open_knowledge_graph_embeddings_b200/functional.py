"""Autograd glue: each hot op of the path as a ``torch.autograd.Function`` over the native kernels.

The reference builds the path out of ATen ops and lets autograd differentiate them
(``openkge/trainer.py:217-234``). Here every forward AND backward is one (or a few) of our own
sm_100a kernels; autograd is only the tape that strings them together, so ``loss.backward()`` and
``optimizer.step()`` keep working exactly as ``Trainer.compute_one_batch`` expects.

Dense-gradient hygiene (matters at N = 10^6 rows): ``LookupAll`` hands out the candidate matrix
``weight[min_size:]`` and the gathered query rows from ONE node, and ``ScoreBCELoss`` /
``ScoreKLLoss`` write dE straight into a ``[min_size + N, D]`` buffer, so the table gradient is
produced in place with no slice-backward / zeros / add passes over 2 GB tensors.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K

PAD_ID = 0        # padding_idx of the embedding tables (openkge/index_mapper.py:14): its row never receives gradient


class TableShadow:
    """fp16 copy of an embedding table, kept on the Parameter (``weight._okge_shadow``): the operand the scoring
    contractions read when the candidate matrix IS the table (Lookup models, 1-vs-all). It is rebuilt (in place, so
    captured CUDA graphs keep valid addresses) whenever the fp32 table changed behind its back, and kept current by the
    fused dE + Adagrad step, which writes the fp16 value of every row it updates (``okge_gemm_adagrad``)."""

    REFRESH_EVERY = 512      # fused updates between two rebuilds (re-derives the power-of-two scale from the live table)

    def __init__(self):
        self.op: Optional[K.F16Operand] = None
        self.version = -1
        self.ptr = 0
        self.dirty = True        # the fp32 table was updated by a native kernel that did not write the fp16 copy
        self.lo_valid = False    # the split-precision (lo) plane matches the table
        self.fused_updates = 0

    def stale(self, weight: torch.Tensor, split: bool) -> bool:
        return (self.op is None or self.dirty or self.version != weight._version or self.ptr != weight.data_ptr()
                or self.op.shape != tuple(weight.shape) or (split and not self.lo_valid))


def table_operand(weight: torch.Tensor, row_start: int = 0, split: bool = False) -> K.F16Operand:
    """fp16 operand of rows [row_start:] of the table ``weight`` (a Parameter), from its shadow copy. ``split``: with the
    lo plane (evaluation)."""
    sh = getattr(weight, "_okge_shadow", None)
    if sh is None:
        sh = TableShadow()
        weight._okge_shadow = sh
    if sh.stale(weight, split):
        w = weight.detach()
        need_lo = split or (sh.op is not None and sh.op.lo is not None)
        out = sh.op if (sh.op is not None and sh.op.shape == tuple(w.shape) and (sh.op.lo is not None or not need_lo)) else None
        sh.op = K.quantize(w, split=need_lo, out=out)
        sh.version, sh.ptr, sh.dirty, sh.lo_valid, sh.fused_updates = weight._version, weight.data_ptr(), False, need_lo, 0
    op = sh.op if split else sh.op.without_lo()
    return op.row_slice(row_start) if row_start else op


def mark_table_updated(weight: torch.Tensor, shadow_current: bool = False) -> None:
    """Called by the optimizers after a native kernel changed ``weight`` in place. ``shadow_current``: the kernel wrote
    the fp16 copy itself (fused step); its lo plane is outdated either way."""
    sh = getattr(weight, "_okge_shadow", None)
    if sh is None:
        return
    sh.lo_valid = False
    if shadow_current:
        sh.fused_updates += 1
        if sh.fused_updates >= TableShadow.REFRESH_EVERY and not torch.cuda.is_current_stream_capturing():
            sh.dirty = True
    else:
        sh.dirty = True


def refresh_table_shadows(module: torch.nn.Module) -> None:
    """Rebuilds (in place) every existing table shadow of ``module``: after parameters were restored behind a captured
    CUDA graph, or to re-derive the scales."""
    for p in module.parameters():
        sh = getattr(p, "_okge_shadow", None)
        if sh is not None and sh.op is not None:
            sh.dirty = True
            table_operand(p, 0, split=sh.op.lo is not None)


class RowsGrad:
    """Sparse gradient of an embedding table whose only use in the step were row lookups: ``grad[ids[i]] += rows[i]`` —
    what ``nn.Embedding(sparse=True)`` (openkge/model.py:390-391) hands to the optimizer as a COO tensor. Consumed by
    ``optim.Adagrad.step`` with the row-wise kernel (torch semantics: only with weight_decay == 0), or materialised."""

    def __init__(self, ids: torch.Tensor, rows: torch.Tensor, shape, skip_id: int):
        self.ids, self.rows, self.shape, self.skip_id = ids, rows, tuple(shape), int(skip_id)

    def merge(self, ids: torch.Tensor, rows: torch.Tensor) -> None:
        self.ids, self.rows = torch.cat([self.ids, ids]), torch.cat([self.rows, rows])

    def materialize(self) -> torch.Tensor:
        gw = torch.zeros(self.shape, dtype=torch.float32, device=self.rows.device)
        K.scatter_add_rows(self.rows, self.ids, gw, self.skip_id)
        return gw

    def adagrad_step(self, param: torch.Tensor, state_sum: torch.Tensor, clr: float, eps: float, weight_decay: float):
        if weight_decay != 0:
            raise RuntimeError("weight_decay option is not compatible with sparse gradients")     # torch.optim.Adagrad's own
        slot_map = getattr(param, "_okge_slot_map", None)
        if slot_map is None or slot_map.numel() != self.shape[0] or slot_map.device != self.rows.device:
            slot_map = torch.full((self.shape[0],), -1, dtype=torch.int32, device=self.rows.device)
            param._okge_slot_map = slot_map
        ids = torch.where(self.ids == self.skip_id, torch.full_like(self.ids, -1), self.ids)
        summed = torch.zeros_like(self.rows)
        K.row_slots_build(ids, slot_map, -1)
        K.row_slots_accumulate(self.rows, ids, slot_map, summed, -1)
        K.adagrad_rows(param.data, state_sum, summed, ids, clr, eps, 0.0, slot_map=slot_map)
        K.row_slots_clear(ids, slot_map, -1)
        mark_table_updated(param)


class GatherRows(torch.autograd.Function):
    """rows = weight[ids]  — nn.Embedding lookup of openkge/model.py:457-458. Dense gradient, or — for a table built with
    ``sparse=True`` (``weight._okge_sparse``) — the sparse form ``RowsGrad`` left on the parameter for the optimizer."""

    @staticmethod
    def forward(ctx, weight: torch.Tensor, ids: torch.Tensor, skip_id: int = -1):
        ids = ids.reshape(-1).to(torch.int32)
        ctx.save_for_backward(ids)
        ctx.shape = weight.shape
        ctx.skip_id = skip_id
        ctx.weight_ref = weight if getattr(weight, "_okge_sparse", False) else None
        return K.gather_rows(weight.detach(), ids)

    @staticmethod
    def backward(ctx, grad):
        (ids,) = ctx.saved_tensors
        weight = ctx.weight_ref
        if weight is not None and weight.grad is None:
            d = getattr(weight, "_okge_deferred", None)
            if d is None:
                weight._okge_deferred = RowsGrad(ids, grad.contiguous(), ctx.shape, ctx.skip_id)
                return None, None, None
            if isinstance(d, RowsGrad):
                d.merge(ids, grad.contiguous())
                return None, None, None
        gw = torch.zeros(ctx.shape, dtype=torch.float32, device=grad.device)
        K.scatter_add_rows(grad.contiguous(), ids, gw, ctx.skip_id)
        return gw, None, None


class AllReduceSum(torch.autograd.Function):
    """y = sum over the ranks of x, the same tensor on every rank afterwards (entity-sharded tables: the rows a rank does
    not own enter as zeros; the rank-local loss sums). backward is the identity: what follows is replicated computation,
    whose gradient with respect to y is already the gradient of the whole job's loss on every rank (``ReplicatedInput``
    sums the rank-local parts where the computation stops being replicated), and dy/dx = 1 for the local summand."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, comm):
        y = x.detach().clone()
        comm.all_reduce(y)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class ReplicatedInput(torch.autograd.Function):
    """Identity on a tensor that is replicated on every rank and feeds rank-local computation (the folded queries in
    front of a rank's block of candidates). backward: the sum over ranks of the rank-local gradients (the all-reduce of
    dQ), so that everything upstream sees the gradient of the whole job's loss."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, comm):
        ctx.comm = comm
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        ctx.comm.all_reduce(g)
        return g, None


class LookupAll(torch.autograd.Function):
    """(E_all, rows) = (weight[min_size:], weight[ids]) from one autograd node.

    E_all is the 1-vs-all candidate operand (``_get_all`` / ``precompute_batch_shared_inputs``,
    openkge/model.py:512-514, 76-77) and aliases the parameter storage (zero copy); rows are the
    known-slot lookups of the batch. backward() reuses the padded buffer that the scoring loss wrote
    dE into (see ``_padded_base``) and scatter-adds the row gradients in place."""

    @staticmethod
    def forward(ctx, weight: torch.Tensor, ids: torch.Tensor, min_size: int):
        ids = ids.reshape(-1).to(torch.int32)
        ctx.save_for_backward(ids)
        ctx.shape = weight.shape
        ctx.min_size = min_size
        ctx.weight_ref = weight          # the Parameter: a deferred gradient is attached to it (DeferredTableGrad)
        ctx.set_materialize_grads(False)  # an output without gradient arrives as None, not as an [N, D] zero tensor
        w = weight.detach()
        rows = K.gather_rows(w, ids)
        e_all = w[min_size:]
        return e_all, rows

    @staticmethod
    def backward(ctx, grad_all: Optional[torch.Tensor], grad_rows: Optional[torch.Tensor]):
        (ids,) = ctx.saved_tensors
        weight = ctx.weight_ref
        pending = _pending_candidate_grads.pop(weight.data_ptr() + ctx.min_size * ctx.shape[1] * 4, None)
        if pending is not None:
            d = DeferredTableGrad(*pending, min_size=ctx.min_size, shape=tuple(ctx.shape))
            if grad_rows is not None and ids.numel():
                d.ids, d.grad_rows = ids, grad_rows.contiguous()
            if grad_all is None and weight.grad is None and getattr(weight, "_okge_deferred", None) is None:
                weight._okge_deferred = d        # consumed by optim.Adagrad.step (fused) or materialised by anyone else
                return None, None, None
            gw = d.materialize()                 # another gradient source exists: fall back to the dense form
            if grad_all is not None:
                gw[ctx.min_size:] += grad_all
            return gw, None, None
        gw = _padded_base(grad_all, ctx.shape, ctx.min_size)
        if gw is None:
            dev = (grad_all if grad_all is not None else grad_rows).device
            gw = torch.zeros(ctx.shape, dtype=torch.float32, device=dev)
            if grad_all is not None:
                gw[ctx.min_size:] = grad_all
        if grad_rows is not None:
            K.scatter_add_rows(grad_rows.contiguous(), ids, gw, PAD_ID)     # padding_idx never receives gradient
        return gw, None, None


def _padded_base(grad_all: Optional[torch.Tensor], shape, min_size: int) -> Optional[torch.Tensor]:
    """If grad_all is rows [min_size:] of a contiguous buffer of the full table shape, return that
    buffer (its first min_size rows are zero by construction)."""
    if grad_all is None:
        return None
    base = grad_all._base
    if base is None or tuple(base.shape) != tuple(shape) or not base.is_contiguous():
        return None
    if grad_all.data_ptr() != base.data_ptr() + min_size * shape[1] * 4 or grad_all.stride(0) != shape[1]:
        return None
    return base


class GatherPool(torch.autograd.Function):
    """pooled = pool_l W[id_rows[row, l]]  — UnigramPoolingRelationEmbedder._encode up to the pooling
    (openkge/model.py:762-774), rows = ids or id_start .. id_start + n."""

    @staticmethod
    def forward(ctx, weight: torch.Tensor, id_rows: torch.Tensor, ids: Optional[torch.Tensor], mode: str,
                id_start: int = 0, n: Optional[int] = None):
        if ids is not None:
            ids = ids.reshape(-1).to(torch.int32)
        w = weight.detach()
        out = K.gather_pool_fwd(w, id_rows, ids, mode, id_start, n)
        ctx.save_for_backward(w, id_rows, ids)
        ctx.mode, ctx.id_start = mode, id_start
        ctx.weight_ref = weight            # the Parameter itself: a compact (slot) gradient is attached to it
        return out

    @staticmethod
    def backward(ctx, grad):
        w, id_rows, ids = ctx.saved_tensors
        weight = ctx.weight_ref
        n_slots = grad.size(0) * id_rows.size(1)
        union = getattr(weight, "_okge_union", None)
        if union is not None:
            # data-parallel step with a touched-row exchange: accumulate into the slots all ranks agreed on (Trainer)
            K.gather_pool_bwd_slots(grad.contiguous(), w, id_rows, ids, ctx.mode, union.write_map, union.buf, ctx.id_start)
            if getattr(weight, "_okge_deferred", None) is None:
                weight._okge_deferred = UnionSlotGrad(union, tuple(w.shape))
            return None, None, None, None, None, None
        # Opt-in (Trainer args["fused_entity_update"]): when the gathered rows name fewer token slots than the table has
        # rows, the gradient goes into a compact slot table and the optimizer applies it without a dense [V, D] gradient
        # (SlotTableGrad). One such node per table and step; a second one (per-block encodes) takes the dense route.
        if (getattr(weight, "_okge_slot_update", False) and 2 * n_slots <= w.size(0) and w.is_contiguous()
                and getattr(weight, "_okge_deferred", None) is None):
            rows = id_rows.index_select(0, ids.long()) if ids is not None else id_rows[ctx.id_start:ctx.id_start + grad.size(0)]
            tok_flat = rows.reshape(-1).contiguous()
            slot_map = getattr(weight, "_okge_slot_map", None)
            if slot_map is None or slot_map.numel() != w.size(0) or slot_map.device != w.device:
                slot_map = torch.full((w.size(0),), -1, dtype=torch.int32, device=w.device)
                weight._okge_slot_map = slot_map
            K.row_slots_build(tok_flat, slot_map, 0)                                  # PAD (0) never receives gradient
            slot_grad = torch.zeros((n_slots, w.size(1)), dtype=torch.float32, device=w.device)
            K.gather_pool_bwd_slots(grad.contiguous(), w, id_rows, ids, ctx.mode, slot_map, slot_grad, ctx.id_start)
            weight._okge_deferred = SlotTableGrad(tok_flat, slot_grad, slot_map, tuple(w.shape))
            return None, None, None, None, None, None
        gw = torch.zeros_like(w)
        K.gather_pool_bwd(grad.contiguous(), w, id_rows, ids, ctx.mode, gw, ctx.id_start)
        return gw, None, None, None, None, None


class UnionSlots:
    """Row numbering shared by all ranks of a data-parallel step: the token rows ANY rank touches in this step, in
    ascending order, get slots base, base + 1, ... of one exchange buffer ``buf`` [cap + 1, D] (all token tables of the
    model share it; row ``cap`` is a dump row that only exists so that a count above the capacity cannot write out of
    bounds). Built by ``union_slots`` from per-rank touch flags with one small all-reduce(max) and a prefix sum; the
    backward pass accumulates straight into ``buf`` (``okge_gather_pool_bwd_slots``), ONE all-reduce of ``buf[:cap]``
    replaces the all-reduce of the dense [V, D] gradients, and the dense optimizer step reads the averaged rows back through
    ``read_map`` (``okge_adagrad_slot_table``)."""

    def __init__(self, write_map: torch.Tensor, read_map: torch.Tensor, buf: torch.Tensor, cap: int):
        self.write_map, self.read_map, self.buf, self.cap = write_map, read_map, buf, int(cap)


def touch_flags(flags: torch.Tensor, id_rows: torch.Tensor, ids: torch.Tensor) -> None:
    """flags[t] = 1 for every token t of the token rows ``id_rows[ids]`` (the PAD token never receives gradient)."""
    tok = id_rows.index_select(0, ids.reshape(-1).long()).reshape(-1).long()
    flags.index_fill_(0, tok, 1)
    flags.narrow(0, PAD_ID, 1).zero_()          # (not ``flags[PAD_ID] = 0``: a host scalar copy cannot be captured)


def union_slots(flags: torch.Tensor, base, cap: int, write_map: torch.Tensor, read_map: torch.Tensor):
    """``flags`` [V] int32 (already all-reduced: 1 = some rank touches the row) -> slot maps (int32 [V], filled in
    place) numbering the touched rows base, base + 1, ... ; returns the next free slot (0-dim int64 device tensor).
    Slots at or above ``cap`` are written to the dump row and read by nobody."""
    pos = torch.cumsum(flags, 0, dtype=torch.int64)
    slot = pos - 1 + base
    on = flags > 0
    write_map.copy_(torch.where(on, torch.clamp(slot, max=cap), -1))
    read_map.copy_(torch.where(on & (slot < cap), slot, -1))
    return pos[-1] + base


class UnionSlotGrad:
    """Gradient of a token table as rows of a ``UnionSlots`` exchange buffer (the averaged rows after the all-reduce)."""

    def __init__(self, union: UnionSlots, shape):
        self.union, self.shape = union, tuple(shape)

    def materialize(self) -> torch.Tensor:
        gw = torch.zeros(self.shape, dtype=torch.float32, device=self.union.buf.device)
        rows = (self.union.read_map >= 0).nonzero().reshape(-1)
        gw[rows] = self.union.buf[self.union.read_map[rows].long()]
        return gw

    def adagrad_step(self, param: torch.Tensor, state_sum: torch.Tensor, clr: float, eps: float, weight_decay: float):
        K.adagrad_slot_table(param.data, state_sum, self.union.read_map, self.union.buf, clr, eps, weight_decay)

    def discard(self) -> None:
        pass


class SlotTableGrad:
    """Gradient of a token table as a compact slot table: ``slot_grad[slot_map[t]]`` is the gradient row of token ``t``
    (slots = positions in the flattened token ids of the gathered rows, ``okge_row_slots_build``), every other row of the
    table has a zero gradient. Consumed by ``optim.Adagrad.step`` like ``DeferredTableGrad``."""

    def __init__(self, tok_flat: torch.Tensor, slot_grad: torch.Tensor, slot_map: torch.Tensor, shape):
        self.tok_flat, self.slot_grad, self.slot_map, self.shape = tok_flat, slot_grad, slot_map, tuple(shape)

    def materialize(self) -> torch.Tensor:
        gw = torch.zeros(self.shape, dtype=torch.float32, device=self.slot_grad.device)
        K.scatter_add_rows(self.slot_grad, self.tok_flat, gw, 0)       # rows of non-winning positions are zero
        K.row_slots_clear(self.tok_flat, self.slot_map, 0)
        return gw

    def adagrad_step(self, param: torch.Tensor, state_sum: torch.Tensor, clr: float, eps: float, weight_decay: float):
        K.adagrad_slot_table(param.data, state_sum, self.slot_map, self.slot_grad, clr, eps, weight_decay)
        K.row_slots_clear(self.tok_flat, self.slot_map, 0)

    def discard(self) -> None:
        """The gradient is dropped without being applied: give the slot map back (all -1)."""
        K.row_slots_clear(self.tok_flat, self.slot_map, 0)


class Dropout(torch.autograd.Function):
    """Counter-based inverted dropout (F.dropout of openkge/model.py:461-470, 783-786); the mask is
    regenerated from (seed, offset) in backward instead of being stored."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, p: float, seed: int, offset: int, step_dev: Optional[torch.Tensor] = None):
        ctx.p, ctx.seed, ctx.offset, ctx.step_dev = p, seed, offset, step_dev
        return K.dropout(x.detach(), p, seed, offset, step_dev)

    @staticmethod
    def backward(ctx, grad):
        return K.dropout(grad.contiguous(), ctx.p, ctx.seed, ctx.offset, ctx.step_dev), None, None, None, None


class LSTMLastState(torch.autograd.Function):
    """Token embeddings -> single-layer LSTM -> hidden state at the last real token of every row
    (LSTMRelationEmbedder._encode_tokens, openkge/model.py:972-980), forward and back-propagation through time.

    ``tok_tm`` [L, n] int32: token ids, TIME-major, so that the rows of one time step are contiguous; ``last_state`` [n]
    int32 in [0, L). Matrix products on the tcgen05 kernel (``gemm_nt``): the input projection of all steps at once, one
    [n, D] x [D, 4D] product per step for the recurrence; the backward mirrors them (dh per step; dW_ih, dW_hh, dX as one
    contraction each over all steps, operands read MN-major in place). The element-wise cell runs in okge_lstm_cell_*.
    Both GEMM operands are raw fp32 (truncated to TF32 by the tensor cores): alpha multiplies the mean shrink back."""

    ALPHA = K.TF32_RAW_OPERAND_SCALE ** 2

    @staticmethod
    def forward(ctx, table, w_ih, w_hh, b_ih, b_hh, tok_tm, last_state):
        L, n = tok_tm.shape
        D = w_hh.size(1)
        dev = table.device
        a = LSTMLastState.ALPHA
        keep = any(ctx.needs_input_grad[:5])              # False under no_grad: nothing is saved, state buffers ping-pong
        w_ih_d, w_hh_d, b_ih_d, b_hh_d = w_ih.detach(), w_hh.detach(), b_ih.detach().contiguous(), b_hh.detach().contiguous()
        flat_tok = tok_tm.reshape(-1)
        X = K.gather_rows(table.detach(), flat_tok)                          # [L * n, D]
        Gx = K.gemm_nt(X, w_ih_d, alpha=a)                                   # [L * n, 4D]
        steps = L if keep else 2                                             # inference: ping-pong state buffers
        C = torch.empty((steps, n, D), dtype=torch.float32, device=dev)
        H = torch.empty((steps, n, D), dtype=torch.float32, device=dev)
        act = torch.empty((L, n, 4 * D), dtype=torch.float32, device=dev) if keep else None
        out = torch.empty((n, D), dtype=torch.float32, device=dev)
        for t in range(L):
            cur, prev = t % steps, (t - 1) % steps
            gh = K.gemm_nt(H[prev], w_hh_d, alpha=a) if t > 0 else None
            K.lstm_cell_fwd(Gx[t * n:(t + 1) * n], gh, b_ih_d, b_hh_d, C[prev] if t > 0 else None, t, last_state,
                            act[t] if keep else None, C[cur], H[cur], out)
        if keep:
            ctx.save_for_backward(w_ih_d, w_hh_d, flat_tok, last_state, X, act, C, H)
            ctx.table_shape = tuple(table.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        w_ih, w_hh, flat_tok, last_state, X, act, C, H = ctx.saved_tensors
        L, n, D4 = act.shape
        D = D4 // 4
        a = LSTMLastState.ALPHA
        g = grad_out.contiguous()
        dG = torch.empty((L, n, D4), dtype=torch.float32, device=g.device)
        dc = torch.zeros((n, D), dtype=torch.float32, device=g.device)
        dh = None
        for t in range(L - 1, -1, -1):
            K.lstm_cell_bwd(act[t], C[t - 1] if t > 0 else None, C[t], g, last_state, t, dh, dc, dG[t])
            if t > 0:
                dh = K.gemm_nt(dG[t], K.ColMajor(w_hh), alpha=a)             # dh_{t-1} = dG_t W_hh
        dGf = dG.view(L * n, D4)
        need = ctx.needs_input_grad
        d_table = d_wih = d_whh = d_b = None
        if need[1]:
            d_wih = K.gemm_nt(K.ColMajor(dGf), K.ColMajor(X), alpha=a)       # dG^T X  [4D, D]
        if need[2]:
            if L > 1:
                d_whh = K.gemm_nt(K.ColMajor(dG[1:].view(-1, D4)), K.ColMajor(H[:L - 1].view(-1, D)), alpha=a)
            else:
                d_whh = torch.zeros_like(w_hh)
        if need[3] or need[4]:
            d_b = dGf.sum(0)
        if need[0]:
            dX = K.gemm_nt(dGf, K.ColMajor(w_ih), alpha=a)                   # [L * n, D]
            d_table = torch.zeros(ctx.table_shape, dtype=torch.float32, device=g.device)
            K.scatter_add_rows(dX, flat_tok, d_table, skip_id=0)             # padding_idx = 0 never receives gradient
        return d_table, d_wih, d_whh, (d_b if need[3] else None), (d_b if need[4] else None), None, None


class SplitRows(torch.autograd.Function):
    """(x[:n], x[n:]) as views; the backward writes the two gradients into one buffer (a plain slice pair would allocate
    two full-size zero tensors and add them)."""

    @staticmethod
    def forward(ctx, x, n: int):
        ctx.n, ctx.shape = int(n), tuple(x.shape)
        return x[:n], x[n:]

    @staticmethod
    def backward(ctx, ga, gb):
        n = ctx.n
        if ga is not None and gb is not None:
            return torch.cat([ga, gb]), None
        g = torch.zeros(ctx.shape, dtype=(ga if ga is not None else gb).dtype, device=(ga if ga is not None else gb).device)
        if ga is not None:
            g[:n] = ga
        if gb is not None:
            g[n:] = gb
        return g, None


class BatchNormRows(torch.autograd.Function):
    """Training-mode ``torch.nn.BatchNorm1d`` over the rows of ``x`` [n, D] (openkge/model.py:463-465, 777-780), with the
    statistics taken per row segment: ``seg`` is an int32 DEVICE tensor of n_seg [begin, end) pairs (None = one segment,
    all rows; rows outside every range are not touched). The reference normalises the po and the sp block of a batch in separate calls; one call with the split as
    data does the same arithmetic with launch shapes that do not depend on the batch. Running statistics of ``bn`` are
    updated in place, segment after segment."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, seg, n_seg, zero_tail=False, dropout=None):
        if bn.momentum is None:
            raise NotImplementedError("cumulative-average batch norm (momentum=None) is not used by the reference")
        track = bn.track_running_stats and bn.running_mean is not None
        y, mean, invstd = K.bn_train_fwd(
            x.detach(), None if gamma is None else gamma.detach(), None if beta is None else beta.detach(),
            bn.running_mean if track else None, bn.running_var if track else None,
            bn.num_batches_tracked if track else None, bn.momentum, bn.eps, seg, n_seg, zero_tail, dropout)
        ctx.save_for_backward(x, gamma, mean, invstd)
        ctx.seg, ctx.n_seg, ctx.zero_tail, ctx.dropout = seg, n_seg, zero_tail, dropout
        return y

    @staticmethod
    def backward(ctx, grad):
        x, gamma, mean, invstd = ctx.saved_tensors
        if ctx.dropout is not None:
            # The kernels can regenerate the mask inside the backward too (okge_bn_train_bwd, drop_p > 0), but its two
            # passes would each pay the Philox rounds and turn issue-bound (measured at 2.5 M x 512: 9.5 ms fused against
            # 2.1 + 4.0 ms for mask pass + backward), so the mask is applied to dy in one separate streaming pass.
            grad = K.dropout(grad.contiguous(), *ctx.dropout)
        dx, dgamma, dbeta = K.bn_train_bwd(grad, x.detach(), None if gamma is None else gamma.detach(), mean, invstd,
                                           ctx.seg, ctx.n_seg, need_dx=ctx.needs_input_grad[0], zero_tail=ctx.zero_tail)
        return (dx, (dgamma if ctx.needs_input_grad[1] else None), (dbeta if ctx.needs_input_grad[2] else None), None, None,
                None, None, None)


def batch_norm_rows(bn: torch.nn.BatchNorm1d, x: torch.Tensor, seg: Optional[torch.Tensor] = None, n_seg: int = 1,
                    segment_rows=None, zero_tail: bool = False, dropout=None) -> torch.Tensor:
    """``bn(x)`` for a 2-D ``x`` through the native kernels. ``segment_rows``: host-side row counts of the segments, when
    the caller knows them, for the reference's error on single-row training batches. ``dropout`` = (p, seed, offset,
    step_dev | None): the inverted dropout that follows the normalisation, fused into the same kernels (training only)."""
    if bn.training:
        sizes = segment_rows if segment_rows is not None else (x.size(0),)
        if any(s == 1 for s in sizes):
            raise ValueError("Expected more than 1 value per channel when training, got input size {}".format([1, x.size(1)]))
        return BatchNormRows.apply(x, bn.weight, bn.bias, bn, seg, n_seg, zero_tail, dropout)
    if torch.is_grad_enabled() and (x.requires_grad or (bn.weight is not None and bn.weight.requires_grad)):
        # eval-mode normalisation inside an autograd graph is off the hot path: plain elementwise ops
        scale = torch.rsqrt(bn.running_var + bn.eps) * (bn.weight if bn.weight is not None else 1.0)
        shift = (bn.bias if bn.bias is not None else 0.0) - bn.running_mean * scale
        return x * scale + shift
    return K.bn_eval_fwd(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


class FoldQuery(torch.autograd.Function):
    """q = fold(kind, a, b): the prefix score of ComplEx / DistMult as one row vector
    (openkge/model.py:206-215, 270-272)."""

    @staticmethod
    def forward(ctx, kind: int, a: torch.Tensor, b: torch.Tensor):
        a, b = a.detach().contiguous(), b.detach().contiguous()
        ctx.kind = kind
        ctx.save_for_backward(a, b)
        return K.fold_query(kind, a, b)

    @staticmethod
    def backward(ctx, gq):
        a, b = ctx.saved_tensors
        ga, gb = K.fold_query_bwd(ctx.kind, a, b, gq.contiguous())
        return None, ga, gb


class FoldQuerySplit(torch.autograd.Function):
    """The folded queries of a whole batch from ONE autograd node: rows [:b_po] are po prefixes (a = obj, b = rel), rows
    [b_po:] sp prefixes (a = subj, b = rel), as AddLossModule orders them (openkge/trainer.py:69-71, 91). Same kernels as
    ``FoldQuery`` on row ranges of the same buffers: no slicing, concatenation or gradient-accumulation nodes in between."""

    @staticmethod
    def forward(ctx, kind_po: int, kind_sp: int, b_po: int, a: torch.Tensor, b: torch.Tensor):
        a, b = a.detach().contiguous(), b.detach().contiguous()
        ctx.kinds, ctx.b_po = (kind_po, kind_sp), int(b_po)
        ctx.save_for_backward(a, b)
        q = torch.empty_like(a)
        for kind, lo, hi in FoldQuerySplit._ranges(kind_po, kind_sp, int(b_po), a.size(0)):
            K.fold_query(kind, a[lo:hi], b[lo:hi], out=q[lo:hi])
        return q

    @staticmethod
    def _ranges(kind_po, kind_sp, b_po, B):
        if kind_po == kind_sp or b_po in (0, B):
            return [(kind_po if b_po else kind_sp, 0, B)]
        return [(kind_po, 0, b_po), (kind_sp, b_po, B)]

    @staticmethod
    def backward(ctx, gq):
        a, b = ctx.saved_tensors
        gq = gq.contiguous()
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        for kind, lo, hi in FoldQuerySplit._ranges(*ctx.kinds, ctx.b_po, a.size(0)):
            K.fold_query_bwd(kind, a[lo:hi], b[lo:hi], gq[lo:hi], out=(ga[lo:hi], gb[lo:hi]))
        return None, None, None, ga, gb


class FoldQueryRows(torch.autograd.Function):
    """Folded queries with the prefix kind of every row in a device tensor (CUDA-graph replay: the po / sp split of the
    batch is data, not a launch shape)."""

    @staticmethod
    def forward(ctx, kinds: torch.Tensor, a: torch.Tensor, b: torch.Tensor):
        a, b = a.detach().contiguous(), b.detach().contiguous()
        ctx.save_for_backward(kinds, a, b)
        return K.fold_query_rows(kinds, a, b)

    @staticmethod
    def backward(ctx, gq):
        kinds, a, b = ctx.saved_tensors
        ga, gb = K.fold_query_rows_bwd(kinds, a, b, gq.contiguous())
        return None, ga, gb


# candidate-operand data_ptr -> (dS panels, q, grad scale): written by the scoring loss backward when the table
# gradient is deferred, consumed by LookupAll.backward of the same table a few autograd nodes later
_pending_candidate_grads = {}


class DeferredTableGrad:
    """Gradient of an embedding table kept in factored form, so that the optimizer can apply it without the
    [N, D] gradient ever being written (``okge_gemm_adagrad``):

        grad[min_size:] = scale * dS^T q          (the 1-vs-all candidate rows, dE of openkge/model.py:206-215)
        grad[ids[i]]   += grad_rows[i]            (the batch's own lookups, embedding_dense_backward of :458)

    ``dS``: fp16 panels of the loss gradient, ``q``: the fp16 query operand of the forward pass. ``materialize()`` builds
    the dense gradient exactly like the unfused backward does; ``adagrad_step`` is torch.optim.Adagrad's dense step
    (utils/optim.py:194-201) on that gradient, fused onto the dE contraction; it also refreshes the table's fp16 copy."""

    def __init__(self, dS, q, scale, dropout=None, *, min_size: int, shape):
        self.dS, self.q, self.scale, self.min_size, self.shape = dS, q, scale, int(min_size), tuple(shape)
        self.dropout = dropout       # (p, seed, offset, step_dev) of the candidate rows' input dropout, or None
        self.ids: Optional[torch.Tensor] = None
        self.grad_rows: Optional[torch.Tensor] = None

    def materialize(self) -> torch.Tensor:
        rows, D = self.shape
        dE = _alloc_dE(rows - self.min_size, D, self.min_size, self.q.device)
        K.gemm_nt(self.dS.T, K.ColMajor(self.q), alpha_dev=self.scale, out=dE, splits=1)
        if self.dropout is not None:
            K.dropout(dE, *self.dropout, out=dE)
        gw = dE._base if self.min_size else dE
        if self.ids is not None:
            K.scatter_add_rows(self.grad_rows, self.ids, gw, PAD_ID)
        return gw

    def adagrad_step(self, param: torch.Tensor, state_sum: torch.Tensor, clr: float, eps: float, weight_decay: float):
        ms, D = self.min_size, self.shape[1]
        data = param.data
        extra = emap = slot_map = None
        if self.ids is not None:
            slot_map = getattr(param, "_okge_slot_map", None)
            if slot_map is None or slot_map.numel() != self.shape[0] or slot_map.device != data.device:
                slot_map = torch.full((self.shape[0],), -1, dtype=torch.int32, device=data.device)
                param._okge_slot_map = slot_map
            extra = torch.zeros((self.ids.numel(), D), dtype=torch.float32, device=data.device)
            K.row_slots_build(self.ids, slot_map, PAD_ID)
            K.row_slots_accumulate(self.grad_rows, self.ids, slot_map, extra, PAD_ID)
            emap = slot_map[ms:]
        sh = getattr(param, "_okge_shadow", None)
        shadow = None
        if sh is not None and sh.op is not None and not sh.dirty and D % 8 == 0:
            shadow = sh.op.without_lo().row_slice(ms)
        K.gemm_adagrad(self.dS.T, K.ColMajor(self.q), data[ms:], state_sum[ms:], clr, eps, weight_decay,
                       alpha_dev=self.scale, extra_map=emap, extra=extra, shadow=shadow, dropout=self.dropout)
        if ms:                                   # the special rows (PAD, UNK) see a zero 1-vs-all gradient
            K.adagrad_slot_rows(data, state_sum, ms, slot_map, extra, clr, eps, weight_decay)
        if slot_map is not None:
            K.row_slots_clear(self.ids, slot_map, PAD_ID)
        mark_table_updated(param, shadow_current=shadow is not None)


def _alloc_dE(N: int, D: int, pad_rows: int, device) -> torch.Tensor:
    """dE buffer as rows [pad_rows:] of a zero-headed [pad_rows + N, D] tensor (see LookupAll)."""
    full = torch.empty((pad_rows + N, D), dtype=torch.float32, device=device)
    if pad_rows:
        full[:pad_rows].zero_()
    return full[pad_rows:]


def _score_backward(dS, q16, e16, e_key: int, grad_scale: torch.Tensor, pad_rows: int, need_q: bool, need_e: bool,
                    defer: bool = False):
    """dQ = g * dS E  and  dE = g * dS^T Q  on the tensor-core kernel (autograd of the mm calls of
    openkge/model.py:206-215). No operand is transposed in memory: the fp16 operands of the forward pass enter
    MN-major (``ColMajor``) and dE reads the dS panels through their transposed view (``dS.T``)."""
    N, D = e16.rows, q16.k
    g = grad_scale.reshape(1).to(torch.float32)
    dQ = dE = None
    if need_q:
        dQ = K.gemm_nt(dS, K.ColMajor(e16), alpha_dev=g)                                  # [B, D], split-K over N
    drop = getattr(e16, "dropout", None)      # the candidate operand was the dropped-out table (AddLossModule.forward)
    if need_e and defer:
        _pending_candidate_grads[e_key] = (dS, q16, g, drop)      # dE = g dS^T q is left to the optimizer (see LookupAll)
    elif need_e:
        dE = _alloc_dE(N, D, pad_rows, q16.device)
        K.gemm_nt(dS.T, K.ColMajor(q16), alpha_dev=g, out=dE, splits=1)                   # [N, D]
        if drop is not None:
            K.dropout(dE, *drop, out=dE)                                                  # d raw = mask / (1 - p) * d dropped
    return dQ, dE


class ScoreBCELoss(torch.autograd.Function):
    """loss_sum = BCEWithLogits(sum)(q E^T, y) with the score matrix kept on chip
    (openkge/trainer.py:91-106); y is CSR positives + (y_base, y_pos) for label smoothing. ``e16``: the fp16 operand of
    ``e`` when the caller already has one (a table's shadow copy); otherwise ``e`` is quantized here, like ``q``."""

    @staticmethod
    def forward(ctx, q, e, pos_ptr, pos_idx, y_base: float, y_pos: float, pad_rows: int = 0, defer_dE: bool = False,
                n_cols_dev: Optional[torch.Tensor] = None, e16: Optional[K.F16Operand] = None):
        need_grad = q.requires_grad or e.requires_grad
        ctx.defer_dE = defer_dE
        q16 = K.quantize(q.detach())
        if e16 is None:
            e16 = K.quantize(e.detach())
        loss, dS = K.score_bce(q16, e16, pos_ptr, pos_idx, y_base, y_pos, want_dS=need_grad, n_cols_dev=n_cols_dev)
        ctx.pad_rows = pad_rows
        if need_grad:
            ctx.operands = (dS, q16, e16, e.data_ptr())
        return loss.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        dS, q16, e16, e_key = ctx.operands
        ctx.operands = None
        dQ, dE = _score_backward(dS, q16, e16, e_key, g, ctx.pad_rows, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                 ctx.defer_dE)
        return dQ, dE, None, None, None, None, None, None, None, None


class ScoreKLLoss(torch.autograd.Function):
    """loss_sum = KLDivLoss(sum)(log_softmax(q E^T, 1), y) (openkge/trainer.py:99-100, 106) from fused
    row log-sum-exp statistics; the gradient pass recomputes the scores tile by tile."""

    @staticmethod
    def forward(ctx, q, e, pos_ptr, pos_idx, pad_rows: int = 0, defer_dE: bool = False, e16: Optional[K.F16Operand] = None,
                shard=None):
        ctx.defer_dE = defer_dE
        q16 = K.quantize(q.detach())
        if e16 is None:
            e16 = K.quantize(e.detach())
        row_lse, pos_score = K.score_lse(q16, e16, pos_ptr, pos_idx)
        npos = (pos_ptr[1:] - pos_ptr[:-1]).to(torch.float32)       # positives of the row on ALL ranks (foreign ones are -1)
        if shard is not None:
            # candidates partitioned over ranks: merge the row statistics (log-sum-exp of log-sum-exps); the returned loss
            # is this rank's summand (rank 0 carries the replicated n_pos * lse term), all-reduced by the caller
            from torch.distributed import ReduceOp
            m = shard.comm.all_reduce(row_lse.clone(), op=ReduceOp.MAX)
            row_lse = m + torch.log(shard.comm.all_reduce(torch.exp(row_lse - m)))
            head = (npos.double() * row_lse.double()).sum() if shard.rank == 0 else 0.0
            loss = head - pos_score.double().sum()
        else:
            loss = (npos.double() * row_lse.double()).sum() - pos_score.double().sum()
        ctx.pad_rows = pad_rows
        ctx.operands = (q16, e16, e.data_ptr())
        ctx.save_for_backward(pos_ptr, pos_idx, row_lse, npos)
        return loss.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        pos_ptr, pos_idx, row_lse, npos = ctx.saved_tensors
        q16, e16, e_key = ctx.operands
        ctx.operands = None
        dS = K.score_softmax_grad(q16, e16, pos_ptr, pos_idx, row_lse, npos)
        dQ, dE = _score_backward(dS, q16, e16, e_key, g, ctx.pad_rows, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                 ctx.defer_dE)
        return dQ, dE, None, None, None, None, None, None


class ScoreMatrix(torch.autograd.Function):
    """scores[B, N] = q E^T materialised — the reference-shaped return value of
    ``_score(prefix=True)`` (openkge/model.py:181-229, 248-278) for callers that want the matrix."""

    @staticmethod
    def forward(ctx, q, e):
        qd, ed = q.detach(), e.detach()
        ctx.save_for_backward(qd, ed)
        return K.score_store(qd, ed)

    @staticmethod
    def backward(ctx, g):
        q, e = ctx.saved_tensors
        g = g.contiguous()
        dQ = dE = None
        if ctx.needs_input_grad[0]:
            dQ = K.gemm_nt(g, K.ColMajor(e), alpha=K.TF32_RAW_OPERAND_SCALE ** 2)    # fp32 operands: the TF32 contraction
        if ctx.needs_input_grad[1]:
            dE = K.gemm_nt(K.ColMajor(g), K.ColMajor(q), alpha=K.TF32_RAW_OPERAND_SCALE ** 2, splits=1)
        return dQ, dE
