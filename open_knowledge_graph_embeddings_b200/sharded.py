"""Entity-sharded training and filtered evaluation over one process per GPU.

The reference has no distributed code (only a non-functional ``nn.DataParallel`` hook,
openkge/trainer.py:143-145); the parity target of this module is the single-device result.

Partition: rank g owns the contiguous block ``[lo_g, hi_g)`` of the N candidate rows of the entity table
together with its optimizer state — neither ever moves. The relation table is replicated. Every rank sees
the same global batch (the sparse collate is cheap and deterministic, so it is replicated instead of being
exchanged) and scores ALL B query rows against ITS candidate block.

Exchanges per training step (NCCL all-reduce over NVLink; gloo in the CPU tests):
  1. X [B, D]   query-side entity rows: every rank contributes the rows it owns, zeros elsewhere  (sum)
  2. dQ [B, D]  partial gradients of the folded queries over the local candidate block           (sum)
  3. loss       one double                                                                        (sum)
Softmax/KL adds the row log-sum-exp merge; evaluation all-reduces the true scores (max) and the int32
(greater, equal) counters (sum) — integer sums are order independent, so ranks are bit-exact for any
number of shards. dE, the entity optimizer step and the relation update need no communication: dE and the
Adagrad step of the local block are ONE kernel (``okge_gemm_adagrad``), the relation gradient is computed
redundantly and identically on every rank.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import kernels as _cuda_kernels
from .dataset import CSRMatrix, RankedAnswers, metrics_from_counts
from .kernels import FOLD_COMPLEX_PO, FOLD_COMPLEX_SP, FOLD_DISTMULT


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous row blocks: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def restrict_csr(ptr: torch.Tensor, idx: torch.Tensor, lo: int, hi: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rows of a CSR matrix restricted to the column block [lo, hi), columns re-based to the block."""
    B = ptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(B, device=ptr.device), (ptr[1:] - ptr[:-1]).long())
    keep = (idx >= lo) & (idx < hi)
    new_ptr = torch.zeros(B + 1, dtype=torch.int32, device=ptr.device)
    new_ptr[1:] = torch.bincount(rows[keep], minlength=B).cumsum(0)
    return new_ptr, (idx[keep] - lo).to(torch.int32)


def local_positions(cols: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """Column -> position inside the block, -1 when another shard owns it."""
    own = (cols >= lo) & (cols < hi)
    return torch.where(own, cols - lo, torch.full_like(cols, -1)).to(torch.int32)


class _Comm:
    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def all_reduce(self, t: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
        if self.on:
            dist.all_reduce(t, op=op, group=self.group)
        return t


class EntityShardedLookupModel:
    """Lookup embedder x {DistMult, ComplEx} scorer with the entity table sharded by rows.

    ``engine`` is the module providing the native ops (default: the CUDA kernels of this package; the CPU
    test-suite injects an oracle-backed engine to exercise the partition / reduction logic under gloo)."""

    def __init__(self, entity_weight_shard: torch.Tensor, relation_weight: torch.Tensor, n_candidates: int, rank: int,
                 world: int, scorer: str = "distmult", offset: int = 2, lr: float = 0.3, eps: float = 1e-8,
                 weight_decay: float = 1e-10, group=None, engine=None):
        self.K = engine if engine is not None else _cuda_kernels
        self.rank, self.world, self.offset, self.N = rank, world, offset, int(n_candidates)
        self.lo, self.hi = shard_bounds(self.N, world, rank)
        assert entity_weight_shard.size(0) == self.hi - self.lo, "shard does not match the partition"
        self.E = entity_weight_shard.contiguous()
        self.R = relation_weight.contiguous()
        self.G_E = torch.zeros_like(self.E)
        self.G_R = torch.zeros_like(self.R)
        self.slot_map = torch.full((self.E.size(0),), -1, dtype=torch.int32, device=self.E.device)
        self.fold_sp = FOLD_COMPLEX_SP if scorer == "complex" else FOLD_DISTMULT
        self.fold_po = FOLD_COMPLEX_PO if scorer == "complex" else FOLD_DISTMULT
        self.lr, self.eps, self.wd = lr, eps, weight_decay
        self.step_count = 0
        self.comm = _Comm(group)

    # ---- query side --------------------------------------------------------------------------
    def _entity_rows(self, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Rows of the (sharded) entity table for global ids: local gather of the owned ones + all-reduce."""
        col = ids.reshape(-1).long() - self.offset
        own = (col >= self.lo) & (col < self.hi)
        own_pos = own.nonzero(as_tuple=False).reshape(-1)
        X = torch.zeros((col.numel(), self.E.size(1)), dtype=torch.float32, device=self.E.device)
        if own_pos.numel():
            X[own_pos] = self.K.gather_rows(self.E, (col[own_pos] - self.lo).to(torch.int32))
        self.comm.all_reduce(X)
        return X, own_pos, (col[own_pos] - self.lo).to(torch.int32)

    def _queries(self, slot_inputs):
        po, sp = slot_inputs
        ent_ids, rel_ids, b_po = [], [], 0
        if po is not None:
            rel_ids.append(po[0].reshape(-1))
            ent_ids.append(po[1].reshape(-1))
            b_po = po[0].numel()
        if sp is not None:
            ent_ids.append(sp[0].reshape(-1))
            rel_ids.append(sp[1].reshape(-1))
        ent_ids, rel_ids = torch.cat(ent_ids), torch.cat(rel_ids).to(torch.int32)
        X, own_pos, own_local = self._entity_rows(ent_ids)
        Rr = self.K.gather_rows(self.R, rel_ids)
        parts = []
        if b_po:
            parts.append(self.K.fold_query(self.fold_po, X[:b_po].contiguous(), Rr[:b_po].contiguous()))
        if b_po < X.size(0):
            parts.append(self.K.fold_query(self.fold_sp, X[b_po:].contiguous(), Rr[b_po:].contiguous()))
        Q = parts[0] if len(parts) == 1 else torch.cat(parts)
        return Q, X, Rr, rel_ids, b_po, own_pos, own_local

    # ---- training ----------------------------------------------------------------------------
    def train_step(self, batch, smoothing: float = 0.0, loss: str = "bce") -> torch.Tensor:
        """One 1-vs-all training step on the global batch; returns the global loss sum (0-dim double)."""
        slot_inputs, normalizer_loss, _, labels, _, _, _ = batch
        K = self.K
        Q, X, Rr, rel_ids, b_po, own_pos, own_local = self._queries(slot_inputs)
        B = Q.size(0)
        ptr_l, idx_l = restrict_csr(labels.ptr, labels.idx, self.lo, self.hi)
        if loss == "bce":
            y_base, y_pos = 0.0, 1.0
            if smoothing > 0:
                y_base, y_pos = (1.0 / self.N) * (1 - smoothing), (1.0 + 1.0 / self.N) * (1 - smoothing)
            loss_part, dS, _ = K.score_bce(Q, self.E, ptr_l, idx_l, y_base, y_pos, want_dST=False)
            loss_sum = self.comm.all_reduce(loss_part.reshape(()).clone())
        else:
            lse_l, pos_l = K.score_lse(Q, self.E, ptr_l, idx_l)
            lse = self._merge_lse(lse_l)
            npos = (labels.ptr[1:] - labels.ptr[:-1]).to(torch.float32)
            pos_sum = self.comm.all_reduce(pos_l.double().sum())
            loss_sum = (npos.double() * lse.double()).sum() - pos_sum
            dS, _ = K.score_softmax_grad(Q, self.E, ptr_l, idx_l, lse, npos, want_dST=False)
        g = 1.0 / float(normalizer_loss)                                   # loss / (B * N), trainer.py:217-221
        dQ = K.gemm_nt(dS, K.ColMajor(self.E), alpha=g * K.TF32_RAW_OPERAND_SCALE)      # E enters MN-major, no transpose pass
        self.comm.all_reduce(dQ)
        dX = torch.empty_like(X)
        dR = torch.empty_like(Rr)
        if b_po:
            dX[:b_po], dR[:b_po] = K.fold_query_bwd(self.fold_po, X[:b_po].contiguous(), Rr[:b_po].contiguous(),
                                                   dQ[:b_po].contiguous())
        if b_po < B:
            dX[b_po:], dR[b_po:] = K.fold_query_bwd(self.fold_sp, X[b_po:].contiguous(), Rr[b_po:].contiguous(),
                                                   dQ[b_po:].contiguous())
        # entity block: dE = g dS^T Q and the Adagrad step in one pass (dE is never written); the lookup gradients of
        # the query rows this rank owns ride along as extra rows. No communication: block and state stay put.
        extra = emap = None
        if own_pos.numel():
            extra = torch.zeros((own_pos.numel(), self.E.size(1)), dtype=torch.float32, device=self.E.device)
            K.row_slots_build(own_local, self.slot_map)
            K.row_slots_accumulate(dX[own_pos].contiguous(), own_local, self.slot_map, extra)
            emap = self.slot_map
        self.step_count += 1
        K.gemm_adagrad(dS.T, K.ColMajor(Q), self.E, self.G_E, self.lr, self.eps, self.wd, alpha=g, extra_map=emap,
                       extra=extra)
        if own_pos.numel():
            K.row_slots_clear(own_local, self.slot_map)
        dRel = torch.zeros_like(self.R)
        K.scatter_add_rows(dR, rel_ids, dRel)
        K.adagrad_dense(self.R, dRel, self.G_R, self.lr, self.eps, self.wd)
        return loss_sum

    def _merge_lse(self, lse_local: torch.Tensor) -> torch.Tensor:
        """log sum_g exp(lse_g): max all-reduce, then sum all-reduce of the rescaled partials."""
        m = self.comm.all_reduce(lse_local.clone(), op=dist.ReduceOp.MAX)
        s = self.comm.all_reduce(torch.exp(lse_local - m))
        return m + torch.log(s)

    # ---- filtered evaluation -----------------------------------------------------------------
    def eval_counts(self, batch):
        """(true_score, greater, equal) of every ranked answer of the global batch, identical on all ranks."""
        slot_inputs, _, _, _, label_ids, filt, _ = batch
        K = self.K
        Q = self._queries(slot_inputs)[0]
        ans: RankedAnswers = label_ids
        dev = Q.device
        n_q = len(ans)
        true = torch.full((n_q,), float("-inf"), dtype=torch.float32, device=dev)
        greater = torch.zeros(n_q, dtype=torch.int32, device=dev)
        equal = torch.zeros(n_q, dtype=torch.int32, device=dev)
        cols = torch.cat([ans.alt_idx, filt.idx]).long()
        own = (cols >= self.lo) & (cols < self.hi)
        uniq, inv = torch.unique(cols[own] - self.lo, return_inverse=True)
        pos = torch.full((cols.numel(),), -1, dtype=torch.int32, device=dev)
        pos[own] = inv.to(torch.int32)
        if uniq.numel():
            sel = K.score_store(Q, K.gather_rows(self.E, uniq.to(torch.int32)))
        else:
            sel = torch.zeros((Q.size(0), 4), dtype=torch.float32, device=dev)
        alt_pos, filt_pos = pos[: ans.alt_idx.numel()].contiguous(), pos[ans.alt_idx.numel():].contiguous()
        K.rank_true_score(sel, ans.ans_row, ans.alt_ptr, alt_pos, true)
        self.comm.all_reduce(true, op=dist.ReduceOp.MAX)                   # alternatives may live on other shards
        K.score_rank(K.gather_rows(Q, ans.ans_row), self.E, true, greater, equal)
        K.rank_filter_correct(sel, ans.ans_row, filt.ptr, filt_pos, true, greater, equal,
                              add_mask_terms=(self.rank == 0))               # the -1e8 fill terms exactly once
        self.comm.all_reduce(greater)
        self.comm.all_reduce(equal)
        return true, greater, equal

    def evaluate_batch(self, batch):
        _, greater, equal = self.eval_counts(batch)
        return metrics_from_counts(greater, equal)
