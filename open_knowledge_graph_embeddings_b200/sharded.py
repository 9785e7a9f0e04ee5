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

from .graphed import CAPTURE_MODE  # noqa: E402
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
    """Rows of a CSR matrix restricted to the column block [lo, hi), columns re-based to the block. The row pointer is
    kept and entries of other blocks become -1 (the label kernels skip negative columns), so that no shape depends on
    device data: no host synchronisation in the sharded step."""
    keep = (idx >= lo) & (idx < hi)
    return ptr, torch.where(keep, idx - lo, torch.full_like(idx, -1)).to(torch.int32)


def local_positions(cols: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """Column -> position inside the block, -1 when another shard owns it."""
    own = (cols >= lo) & (cols < hi)
    return torch.where(own, cols - lo, torch.full_like(cols, -1)).to(torch.int32)


class EntityShard:
    """This rank's block [lo, hi) of the real entity rows (``RelationEmbedder.shard_entities``)."""

    def __init__(self, lo: int, hi: int, rank: int, world: int, comm: "_Comm"):
        self.lo, self.hi, self.rank, self.world, self.comm = int(lo), int(hi), int(rank), int(world), comm


class _Comm:
    """``group="local"``: no communication at all (a single-rank model inside a multi-rank job, e.g. the reference run of
    the on-hardware parity check)."""

    def __init__(self, group=None):
        self.group = None if group == "local" else group
        self.on = group != "local" and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def all_reduce(self, t: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
        if self.on:
            dist.all_reduce(t, op=op, group=self.group)
        return t


def sharded_rank_counts(K, comm: "_Comm", Q: torch.Tensor, E_block: torch.Tensor, lo: int, hi: int, rank: int,
                        ans: RankedAnswers, filt: CSRMatrix, e16=None, split: bool = True):
    """Filtered rank counts of ``OneToNMentionRelationDataset.compute_metrics`` (openkge/dataset.py:423-445) when the
    candidate rows [lo, hi) live on this rank. The true score of an answer is the max over its alternative mentions,
    which may sit on other shards (max all-reduce); (greater, equal) are int32 counts over the local block, corrected
    for the local filter columns, then summed over ranks: integer sums, so the result is bit-exact for any partition.
    ``e16``: fp16 operand of ``E_block`` if the caller keeps one; ``split``: split-precision scores (fp32-grade)."""
    dev = Q.device
    n_q = len(ans)
    true = torch.full((n_q,), float("-inf"), dtype=torch.float32, device=dev)
    greater = torch.zeros(n_q, dtype=torch.int32, device=dev)
    equal = torch.zeros(n_q, dtype=torch.int32, device=dev)
    q16 = K.quantize(Q, split=split)
    if e16 is None and hi > lo:
        e16 = K.quantize(E_block, split=split)
    # columns owned by this shard keep their position in `sel`, the others are marked -1; foreign columns are gathered
    # from a clamped row (never read) so that no shape depends on device data (no host synchronisation)
    cols = torch.cat([ans.alt_idx, filt.idx]).long()
    own = (cols >= lo) & (cols < hi)
    n_alt = ans.alt_idx.numel()
    if cols.numel() and hi > lo:
        local = torch.where(own, cols - lo, torch.zeros_like(cols)).to(torch.int32)
        sel = K.score_store(q16, K.gather_rows_f16(e16, local), split=split)
    else:
        sel = torch.zeros((Q.size(0), 4), dtype=torch.float32, device=dev)
    pos = torch.where(own, torch.arange(cols.numel(), device=dev), torch.full_like(cols, -1)).to(torch.int32)
    alt_pos, filt_pos = pos[:n_alt].contiguous(), pos[n_alt:].contiguous()
    K.rank_true_score(sel, ans.ans_row, ans.alt_ptr, alt_pos, true)
    comm.all_reduce(true, op=dist.ReduceOp.MAX)                        # alternatives may live on other shards
    if hi > lo:
        K.score_rank(K.gather_rows_f16(q16, ans.ans_row), e16, true, greater, equal, split=split)
    K.rank_filter_correct(sel, ans.ans_row, filt.ptr, filt_pos, true, greater, equal,
                          add_mask_terms=(rank == 0))                  # the -1e8 fill terms exactly once
    comm.all_reduce(greater)
    comm.all_reduce(equal)
    return true, greater, equal


def merge_lookup_shards(shards) -> Tuple[dict, dict, int]:
    """Reassembles the per-rank files of ``EntityShardedLookupModel.state_dict_shard`` into the reference's tensors:
    (model state dict with ``entity_embedding.weight`` [N + offset, D] and ``relation_embedding.weight`` — loadable with
    ``Models.Lookup*RelationModel.load_state_dict(strict=True)`` —, the Adagrad ``sum`` accumulators under the same keys,
    training_steps). Works on CPU tensors; the replicated tensors are taken from rank 0."""
    shards = sorted(shards, key=lambda d: d["rank"])
    world = shards[0]["world"]
    if [d["rank"] for d in shards] != list(range(world)) or any(d["world"] != world for d in shards):
        raise ValueError("need exactly one shard of every rank of the same run")
    if any(a["hi"] != b["lo"] for a, b in zip(shards, shards[1:])) or shards[0]["lo"] != 0 or \
            shards[-1]["hi"] != shards[0]["n_candidates"]:
        raise ValueError("shards do not tile the candidate range")
    first = shards[0]
    cpu = lambda t: t.detach().to("cpu")
    state = {"entity_embedding.weight": torch.cat([cpu(first["entity_embedding.special_rows"])] +
                                                  [cpu(d["entity_embedding.weight"]) for d in shards]),
             "relation_embedding.weight": cpu(first["relation_embedding.weight"])}
    sums = {"entity_embedding.weight": torch.cat([cpu(first["entity_embedding.special_rows/sum"])] +
                                                 [cpu(d["entity_embedding.weight/sum"]) for d in shards]),
            "relation_embedding.weight": cpu(first["relation_embedding.weight/sum"])}
    return state, sums, int(first["training_steps"])


class EntityShardedLookupModel:
    """Lookup embedder x {DistMult, ComplEx} scorer with the entity table sharded by rows.

    ``engine`` is the module providing the native ops (default: the CUDA kernels of this package; the CPU
    test-suite injects an oracle-backed engine to exercise the partition / reduction logic under gloo)."""

    def __init__(self, entity_weight_shard: torch.Tensor, relation_weight: torch.Tensor, n_candidates: int, rank: int,
                 world: int, scorer: str = "distmult", offset: int = 2, lr: float = 0.3, eps: float = 1e-8,
                 weight_decay: float = 1e-10, group=None, engine=None, special_rows: Optional[torch.Tensor] = None,
                 fused_update_max_rows: Optional[int] = None):
        self.K = engine if engine is not None else _cuda_kernels
        # The block's gradient is applied by the fused dE + Adagrad kernel at every batch size (fp16 operands, B200: 0.66 ms
        # at 4,096 rows x 125 k-row block against 0.43 + 0.23 + 0.21 ms for the dE contraction, the dense Adagrad kernel and
        # the re-quantised scoring operand; 0.78 against 1.32 ms at 2,048 x 250 k). ``fused_update_max_rows`` is a test hook:
        # batches with more rows than it take the unfused route (0 = always unfused).
        self.fused_update_max_rows = None if fused_update_max_rows is None else int(fused_update_max_rows)
        self.rank, self.world, self.offset, self.N = rank, world, offset, int(n_candidates)
        self.lo, self.hi = shard_bounds(self.N, world, rank)
        assert entity_weight_shard.size(0) == self.hi - self.lo, "shard does not match the partition"
        self.E = entity_weight_shard.contiguous()
        self.R = relation_weight.contiguous()
        self.G_E = torch.zeros_like(self.E)
        self.G_R = torch.zeros_like(self.R)
        # rows 0 .. offset-1 of the reference table (PAD, UNK): never candidates and never looked up, but the reference's
        # dense Adagrad decays them with every step (weight decay), so they are carried along, replicated, for
        # checkpoints that are interchangeable with the single-device model
        self.special = (special_rows.to(self.E.device, torch.float32).contiguous().clone() if special_rows is not None
                        else torch.zeros((offset, self.E.size(1)), dtype=torch.float32, device=self.E.device))
        self.G_special = torch.zeros_like(self.special)
        self.scorer = scorer
        self.row_kinds: Optional[torch.Tensor] = None       # int32 [B] fold kind per row, set by GraphedShardedStep
        self.slot_map = torch.full((self.E.size(0),), -1, dtype=torch.int32, device=self.E.device)
        self.fold_sp = FOLD_COMPLEX_SP if scorer == "complex" else FOLD_DISTMULT
        self.fold_po = FOLD_COMPLEX_PO if scorer == "complex" else FOLD_DISTMULT
        self.lr, self.eps, self.wd = lr, eps, weight_decay
        self.step_count = 0
        self.comm = _Comm(group)
        self._e16 = None          # fp16 operand of the block (kept current by the fused update), see _block_operand
        self._e16_dirty = True
        self._e16_lo_valid = False

    def _block_operand(self, split: bool = False):
        """fp16 operand of this rank's block: rebuilt in place after an unfused update or a checkpoint load, otherwise
        maintained by ``okge_gemm_adagrad``; the lo plane (evaluation) is refreshed on demand."""
        if self.E.size(0) == 0:
            return None
        if self._e16 is None or self._e16_dirty or (split and not self._e16_lo_valid):
            need_lo = split or (self._e16 is not None and getattr(self._e16, "lo", None) is not None)
            self._e16 = self.K.quantize(self.E, split=need_lo, out=self._e16 if (self._e16 is not None and (
                not need_lo or getattr(self._e16, "lo", None) is not None)) else None)
            self._e16_dirty, self._e16_lo_valid = False, need_lo
        if split or not hasattr(self._e16, "without_lo"):
            return self._e16
        return self._e16.without_lo()

    # ---- query side --------------------------------------------------------------------------
    def _entity_rows(self, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rows of the (sharded) entity table for global ids: every rank gathers the rows it owns (zeros elsewhere), one
        all-reduce assembles them. Returns (X [B, D], local [B] int32 = row in this shard or -1). Shapes never depend on
        device data, so the step has no host synchronisation."""
        col = ids.reshape(-1).long() - self.offset
        own = (col >= self.lo) & (col < self.hi)
        local = torch.where(own, col - self.lo, torch.full_like(col, -1)).to(torch.int32)
        if self.E.size(0):
            X = self.K.gather_rows(self.E, local.clamp(min=0)) * own.unsqueeze(1).to(self.E.dtype)
        else:
            X = torch.zeros((col.numel(), self.E.size(1)), dtype=torch.float32, device=self.E.device)
        self.comm.all_reduce(X)
        return X, local

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
        X, local = self._entity_rows(ent_ids)
        Rr = self.K.gather_rows(self.R, rel_ids)
        if self.row_kinds is not None:                      # graph replay: the po / sp kind of every row is data
            return self.K.fold_query_rows(self.row_kinds, X, Rr), X, Rr, rel_ids, b_po, local
        if self.fold_po == self.fold_sp:                    # DistMult: one fold for the whole batch
            return self.K.fold_query(self.fold_sp, X, Rr), X, Rr, rel_ids, b_po, local
        parts = []
        if b_po:
            parts.append(self.K.fold_query(self.fold_po, X[:b_po].contiguous(), Rr[:b_po].contiguous()))
        if b_po < X.size(0):
            parts.append(self.K.fold_query(self.fold_sp, X[b_po:].contiguous(), Rr[b_po:].contiguous()))
        Q = parts[0] if len(parts) == 1 else torch.cat(parts)
        return Q, X, Rr, rel_ids, b_po, local

    # ---- training ----------------------------------------------------------------------------
    def train_step(self, batch, smoothing: float = 0.0, loss: str = "bce") -> torch.Tensor:
        """One 1-vs-all training step on the global batch; returns the global loss sum (0-dim double)."""
        slot_inputs, normalizer_loss, _, labels, _, _, _ = batch
        K = self.K
        Q, X, Rr, rel_ids, b_po, local = self._queries(slot_inputs)
        B = Q.size(0)
        ptr_l, idx_l = restrict_csr(labels.ptr, labels.idx, self.lo, self.hi)
        q16, e16 = K.quantize(Q), self._block_operand()
        if loss == "bce":
            y_base, y_pos = 0.0, 1.0
            if smoothing > 0:
                y_base, y_pos = (1.0 / self.N) * (1 - smoothing), (1.0 + 1.0 / self.N) * (1 - smoothing)
            loss_part, dS = K.score_bce(q16, e16, ptr_l, idx_l, y_base, y_pos)
            loss_sum = self.comm.all_reduce(loss_part.reshape(()).clone())
        else:
            lse_l, pos_l = K.score_lse(q16, e16, ptr_l, idx_l)
            lse = self._merge_lse(lse_l)
            npos = (labels.ptr[1:] - labels.ptr[:-1]).to(torch.float32)
            pos_sum = self.comm.all_reduce(pos_l.double().sum())
            loss_sum = (npos.double() * lse.double()).sum() - pos_sum
            dS = K.score_softmax_grad(q16, e16, ptr_l, idx_l, lse, npos)
        g = 1.0 / float(normalizer_loss)                                   # loss / (B * N), trainer.py:217-221
        dQ = K.gemm_nt(dS, K.ColMajor(e16), alpha=g)                       # E enters MN-major, no transpose pass
        self.comm.all_reduce(dQ)
        if self.row_kinds is not None:
            dX, dR = K.fold_query_rows_bwd(self.row_kinds, X, Rr, dQ.contiguous())
        elif self.fold_po == self.fold_sp:
            dX, dR = K.fold_query_bwd(self.fold_sp, X, Rr, dQ.contiguous())
        else:
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
        # (rows this rank does not own carry local id -1 and are skipped by the slot kernels)
        self.step_count += 1
        if self.fused_update_max_rows is None or B <= self.fused_update_max_rows:
            extra = torch.zeros_like(dX)
            K.row_slots_build(local, self.slot_map, -1)
            K.row_slots_accumulate(dX, local, self.slot_map, extra, -1)
            K.gemm_adagrad(dS.T, K.ColMajor(q16), self.E, self.G_E, self.lr, self.eps, self.wd, alpha=g, extra_map=self.slot_map,
                           extra=extra, shadow=e16 if self.E.size(1) % 8 == 0 else None)
            K.row_slots_clear(local, self.slot_map, -1)
            self._e16_dirty = self.E.size(1) % 8 != 0
        else:
            dE = K.gemm_nt(dS.T, K.ColMajor(q16), alpha=g)                # [rows of the block, D]
            K.scatter_add_rows(dX.contiguous(), local, dE, -1)            # lookup gradients of the rows this rank owns
            K.adagrad_dense(self.E, dE, self.G_E, self.lr, self.eps, self.wd)
            self._e16_dirty = True
        self._e16_lo_valid = False
        dRel = torch.zeros_like(self.R)
        K.scatter_add_rows(dR, rel_ids, dRel)
        K.adagrad_dense(self.R, dRel, self.G_R, self.lr, self.eps, self.wd)
        if self.special.numel():
            K.adagrad_dense(self.special, torch.zeros_like(self.special), self.G_special, self.lr, self.eps, self.wd)
        return loss_sum

    # ---- checkpoints -------------------------------------------------------------------------
    # The reference saves one file: {"state_dict": model.state_dict(), "optimizer_state_dict": [...], "training_steps": ...}
    # (openkge/trainer.py:608-620) with the whole entity table in `entity_embedding.weight`. A sharded run writes one file
    # per rank (its row block + Adagrad accumulators; the replicated tensors ride along in every file) and
    # `merge_lookup_shards` reassembles the reference's tensors; `from_reference_state_dict` goes the other way, for any
    # number of ranks.
    def state_dict_shard(self) -> dict:
        return {"format": "okge_b200.entity_shard.v1", "rank": self.rank, "world": self.world, "n_candidates": self.N,
                "offset": self.offset, "lo": self.lo, "hi": self.hi, "scorer": self.scorer, "training_steps": self.step_count,
                "entity_embedding.weight": self.E, "entity_embedding.weight/sum": self.G_E,
                "entity_embedding.special_rows": self.special, "entity_embedding.special_rows/sum": self.G_special,
                "relation_embedding.weight": self.R, "relation_embedding.weight/sum": self.G_R}

    def load_state_dict_shard(self, shard: dict) -> None:
        if (shard["world"], shard["rank"], shard["n_candidates"]) != (self.world, self.rank, self.N):
            raise ValueError("shard was written for a different partition; merge the shards and use "
                             "from_reference_state_dict to re-partition")
        for key, dst in (("entity_embedding.weight", self.E), ("entity_embedding.weight/sum", self.G_E),
                         ("entity_embedding.special_rows", self.special), ("entity_embedding.special_rows/sum", self.G_special),
                         ("relation_embedding.weight", self.R), ("relation_embedding.weight/sum", self.G_R)):
            dst.copy_(shard[key])
        self.step_count = int(shard["training_steps"])
        self._e16_dirty = True

    @classmethod
    def from_reference_state_dict(cls, state_dict: dict, rank: int, world: int, device, scorer: str = "distmult",
                                  adagrad_sums: Optional[dict] = None, training_steps: int = 0, **kwargs):
        """The rank's shard of a single-device checkpoint: ``state_dict`` holds the reference's keys
        ``entity_embedding.weight`` [N + offset, D] and ``relation_embedding.weight``; ``adagrad_sums`` optionally the
        Adagrad accumulators under the same keys (the ``sum`` entries of the reference's optimizer state)."""
        offset = kwargs.get("offset", 2)
        W = state_dict["entity_embedding.weight"]
        N = W.size(0) - offset
        lo, hi = shard_bounds(N, world, rank)
        model = cls(W[offset + lo:offset + hi].to(device, torch.float32).clone(),
                    state_dict["relation_embedding.weight"].to(device, torch.float32).clone(), N, rank, world,
                    scorer=scorer, special_rows=W[:offset], **kwargs)
        if adagrad_sums is not None:
            G = adagrad_sums["entity_embedding.weight"]
            model.G_E.copy_(G[offset + lo:offset + hi])
            model.G_special.copy_(G[:offset])
            model.G_R.copy_(adagrad_sums["relation_embedding.weight"])
        model.step_count = int(training_steps)
        return model

    def _merge_lse(self, lse_local: torch.Tensor) -> torch.Tensor:
        """log sum_g exp(lse_g): max all-reduce, then sum all-reduce of the rescaled partials."""
        m = self.comm.all_reduce(lse_local.clone(), op=dist.ReduceOp.MAX)
        s = self.comm.all_reduce(torch.exp(lse_local - m))
        return m + torch.log(s)

    # ---- filtered evaluation -----------------------------------------------------------------
    def eval_counts(self, batch):
        """(true_score, greater, equal) of every ranked answer of the global batch, identical on all ranks."""
        slot_inputs, _, _, _, label_ids, filt, _ = batch
        Q = self._queries(slot_inputs)[0]
        return sharded_rank_counts(self.K, self.comm, Q, self.E, self.lo, self.hi, self.rank, label_ids, filt,
                                   e16=self._block_operand(split=True), split=True)

    def evaluate_batch(self, batch):
        _, greater, equal = self.eval_counts(batch)
        return metrics_from_counts(greater, equal)


class GraphedShardedStep:
    """CUDA-graph replay of ``EntityShardedLookupModel.train_step``: kernels AND the NCCL all-reduces of one step are
    captured once (every rank captures the same sequence) and replayed per batch from static device buffers, like
    ``graphed.GraphedTrainStep`` on one GPU. The step has no shape that depends on device data (foreign rows / columns are
    marked -1, never compacted) and no host synchronisation, which is what makes it capturable.

    ``rows``: prefix rows of the GLOBAL batch; ``max_positives``: capacity of the CSR column buffer."""

    def __init__(self, model: "EntityShardedLookupModel", rows: int, max_positives: int, example_batch,
                 smoothing: float = 0.0, loss: str = "bce", preserve_state: bool = True):
        self.model, self.rows, self.capacity = model, int(rows), int(max_positives)
        dev = model.E.device
        self.ent = torch.zeros((rows, 1), dtype=torch.int32, device=dev)
        self.rel = torch.zeros((rows, 1), dtype=torch.int32, device=dev)
        self.ptr = torch.zeros(rows + 1, dtype=torch.int32, device=dev)
        self.idx = torch.full((self.capacity,), -1, dtype=torch.int32, device=dev)
        self.asymmetric = model.fold_po != model.fold_sp
        self.kinds = torch.full((rows,), int(model.fold_sp), dtype=torch.int32, device=dev)
        self.normalizer_loss = rows * model.N
        self.static_batch = ([None, (self.ent, self.rel)], self.normalizer_loss, 0.0,
                             CSRMatrix(self.ptr, self.idx, (rows, model.N)), None, None, None)
        self.smoothing, self.loss_kind = smoothing, loss
        # the warm-up and capture runs below are real training steps on `example_batch`: unless the caller wants them
        # (preserve_state=False), weights, optimizer state and the step counter are put back afterwards, in place (the
        # graph holds the addresses), so that creating the graphed step does not change the training trajectory
        state = (model.E, model.R, model.G_E, model.G_R, model.special, model.G_special)
        snapshot = ([t.detach().clone() for t in state], model.step_count) if preserve_state else None
        self.load(example_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=CAPTURE_MODE):
            self.loss = self._eager()
        if snapshot is not None:
            for dst, src in zip(state, snapshot[0]):
                dst.copy_(src)
            model.step_count = snapshot[1]
            model._e16_dirty = True
            model._block_operand()                   # the fp16 copy of the block follows, same buffer
            torch.cuda.synchronize()

    def _eager(self):
        model = self.model
        saved = model.row_kinds
        model.row_kinds = self.kinds if self.asymmetric else None
        try:
            return model.train_step(self.static_batch, smoothing=self.smoothing, loss=self.loss_kind)
        finally:
            model.row_kinds = saved

    def load(self, batch) -> None:
        slot_inputs, normalizer_loss, _, labels, _, _, _ = batch
        po, sp = slot_inputs
        ent = [t for t in ((po[1] if po is not None else None), (sp[0] if sp is not None else None)) if t is not None]
        rel = [t for t in ((po[0] if po is not None else None), (sp[1] if sp is not None else None)) if t is not None]
        ent = ent[0] if len(ent) == 1 else torch.cat(ent)
        rel = rel[0] if len(rel) == 1 else torch.cat(rel)
        nnz = labels.idx.numel()
        if ent.numel() != self.rows or normalizer_loss != self.normalizer_loss or nnz > self.capacity:
            raise ValueError(f"graphed step was captured for {self.rows} rows and at most {self.capacity} positives")
        self.ent.copy_(ent.reshape(-1, 1), non_blocking=True)
        self.rel.copy_(rel.reshape(-1, 1), non_blocking=True)
        if self.asymmetric:
            b_po = 0 if po is None else po[0].numel()
            if b_po != getattr(self, "_b_po", None):
                self.kinds[:b_po] = int(self.model.fold_po)
                self.kinds[b_po:] = int(self.model.fold_sp)
                self._b_po = b_po
        self.ptr.copy_(labels.ptr, non_blocking=True)
        self.idx[:nnz].copy_(labels.idx, non_blocking=True)

    def __call__(self, batch) -> torch.Tensor:
        """One training step on ``batch``; returns the global loss sum (device tensor, valid until the next call)."""
        self.load(batch)
        self.graph.replay()
        self.model.step_count += 1
        return self.loss


class CandidateShardedUnigramModel:
    """UnigramPooling embedder x {ComplEx, DistMult} scorer (openkge/model.py:716-798, 1016-1019) over one process per
    GPU, for the OLPBench-shaped configs.

    Replicated: both token tables, the token-id rows, batch-norm parameters and all optimizer state (0.5 GB at C4).
    Partitioned: the CANDIDATE list of a step — every real mention in 1-vs-all training / evaluation (2.5 M pooled
    rows at C5), or the batch-shared candidate ids (openkge/dataset.py:813-868) — rank g encodes and scores positions
    [lo_g, hi_g). Every rank encodes all B query rows itself (the token tables are local), so no query exchange.

    Exchanges per training step: batch-norm statistics of the candidate encode (2 x D sums, forward and backward),
    dQ [B, D] partials, the loss, and the token-table / batch-norm gradients of the candidate side (the query side is
    computed redundantly and identically everywhere and is added once, after the all-reduce). Evaluation: true scores
    (max) and int32 counts (sum), see ``sharded_rank_counts``. Dropout is not supported here (p = 0)."""

    BN_EPS, BN_MOMENTUM = 1e-5, 0.1

    def __init__(self, params: dict, n_candidates: int, rank: int, world: int, scorer: str = "complex", pool: str = "sum",
                 offset: int = 2, lr: float = 0.1, eps: float = 1e-8, weight_decay: float = 1e-10, group=None,
                 engine=None):
        self.K = engine if engine is not None else _cuda_kernels
        self.rank, self.world, self.offset, self.N = rank, world, offset, int(n_candidates)
        self.pool = pool
        self.p = {k: v for k, v in params.items()}           # reference state-dict keys (openkge/model.py:597-634)
        self.rows = {w: self.p[f"{w}_token_ids"].to(torch.int32).contiguous() for w in ("entity", "relation")}
        self.batchnorm = "entity_batchnorm.weight" in self.p
        self.trainable = [k for k in self.p if k.endswith("embedding.weight") or k.endswith("batchnorm.weight")
                          or k.endswith("batchnorm.bias")]
        self.state_sum = {k: torch.zeros_like(self.p[k]) for k in self.trainable}
        self.fold_sp = FOLD_COMPLEX_SP if scorer == "complex" else FOLD_DISTMULT
        self.fold_po = FOLD_COMPLEX_PO if scorer == "complex" else FOLD_DISTMULT
        self.lr, self.eps, self.wd = lr, eps, weight_decay
        self.comm = _Comm(group)
        self._eval_cache = None

    # ---- encoders ----------------------------------------------------------------------------
    def state_dict(self) -> dict:
        """Every tensor is replicated, under the reference's own keys (token tables, token-id rows, batch-norm tensors):
        rank 0 saves this as the ``state_dict`` of a reference checkpoint (openkge/trainer.py:608-620)."""
        return {k: v.detach().clone() for k, v in self.p.items()}

    def optimizer_sums(self) -> dict:
        """Adagrad accumulators of the trainable tensors, keyed like ``state_dict``."""
        return {k: v.detach().clone() for k, v in self.state_sum.items()}

    def _encode(self, which: str, ids: torch.Tensor, training: bool, sync: bool, n_total: Optional[int] = None):
        """pool -> batch norm (openkge/model.py:762-780). ``sync``: the rows are one rank's share of a partitioned
        batch of ``n_total`` rows, statistics are summed over the ranks. Returns (y, cache for the backward pass)."""
        ids = ids.reshape(-1).to(torch.int32)
        x = self.K.gather_pool_fwd(self.p[f"{which}_embedding.weight"], self.rows[which], ids, self.pool)
        if not self.batchnorm:
            return x, (which, ids, None)
        g, b = self.p[f"{which}_batchnorm.weight"], self.p[f"{which}_batchnorm.bias"]
        rm, rv = self.p[f"{which}_batchnorm.running_mean"], self.p[f"{which}_batchnorm.running_var"]
        if not training:
            inv = torch.rsqrt(rv + self.BN_EPS)
            return self.K.bn_normalize(x, rm, inv, g, b), (which, ids, None)
        # three phases on the native kernels: fp64 column sums of the local rows -> all-reduce -> normalise
        stats = self.K.bn_col_sums(x)                                    # [2, D] fp64: sum x, sum x^2
        if sync:
            self.comm.all_reduce(stats)
        n = float(n_total if (sync and n_total is not None) else x.size(0))     # known on the host: no device read
        mean = stats[0] / n
        var = (stats[1] / n - mean * mean).clamp_min(0.0)                # biased, used for normalisation
        mean_f = mean.float()
        inv = (1.0 / torch.sqrt(var.float() + self.BN_EPS))
        y = self.K.bn_normalize(x, mean_f, inv, g, b)
        m = self.BN_MOMENTUM                                             # running stats: unbiased variance (torch BatchNorm1d)
        self.p[f"{which}_batchnorm.running_mean"] = ((1 - m) * rm + m * mean_f)
        self.p[f"{which}_batchnorm.running_var"] = ((1 - m) * rv + m * (var * n / max(n - 1.0, 1.0)).float())
        self.p[f"{which}_batchnorm.num_batches_tracked"] = self.p[f"{which}_batchnorm.num_batches_tracked"] + 1
        return y, (which, ids, (x, mean_f, inv, n, sync))

    def _encode_backward(self, grad_y: torch.Tensor, cache, grads: dict):
        which, ids, bn = cache
        g = grad_y
        if bn is not None:
            x, mean, inv, n, sync = bn
            gamma = self.p[f"{which}_batchnorm.weight"]
            sums = self.K.bn_col_sums(grad_y, x, mean, inv)              # [2, D] fp64: d beta, d gamma
            if sync:
                self.comm.all_reduce(sums)
            grads[f"{which}_batchnorm.bias"] += sums[0].float()
            grads[f"{which}_batchnorm.weight"] += sums[1].float()
            # dx = gamma * inv * (dy - sum dy / n - xhat * sum dy xhat / n)
            g = self.K.bn_normalize_bwd(grad_y, x, mean, inv, (sums / n).float().contiguous(), gamma)
        key = f"{which}_embedding.weight"
        self.K.gather_pool_bwd(g.contiguous(), self.p[key], self.rows[which], ids, self.pool, grads[key])

    def _candidate_ids(self, shared) -> torch.Tensor:
        dev = self.p["entity_embedding.weight"].device
        if isinstance(shared, torch.Tensor):
            return shared.reshape(-1).to(device=dev, dtype=torch.int32)
        return torch.arange(self.offset, self.offset + self.N, dtype=torch.int32, device=dev)

    def _queries(self, slot_inputs, training: bool):
        """Query rows in the reference's call order (openkge/trainer.py:69-87): po rel, po obj, sp subj, sp rel."""
        po, sp = slot_inputs
        K = self.K
        parts, tape = [], []
        if po is not None:
            rel, c_rel = self._encode("relation", po[0], training, False)
            obj, c_obj = self._encode("entity", po[1], training, False)
            parts.append(K.fold_query(self.fold_po, obj.contiguous(), rel.contiguous()))
            tape.append((self.fold_po, obj, rel, c_obj, c_rel))
        if sp is not None:
            subj, c_subj = self._encode("entity", sp[0], training, False)
            rel, c_rel = self._encode("relation", sp[1], training, False)
            parts.append(K.fold_query(self.fold_sp, subj.contiguous(), rel.contiguous()))
            tape.append((self.fold_sp, subj, rel, c_subj, c_rel))
        Q = parts[0] if len(parts) == 1 else torch.cat(parts)
        return Q, tape

    # ---- training ----------------------------------------------------------------------------
    def train_step(self, batch, smoothing: float = 0.0) -> torch.Tensor:
        """One BCE training step on the global batch (1-vs-all or batch-shared candidates); returns the global loss sum."""
        slot_inputs, normalizer_loss, _, labels, _, _, shared = batch
        K = self.K
        cand = self._candidate_ids(shared)
        n_c = cand.numel()
        lo, hi = shard_bounds(n_c, self.world, self.rank)
        self._eval_cache = None
        # candidates first (entity batch norm call #1, statistics over ALL candidates), then the query rows
        E, c_cand = self._encode("entity", cand[lo:hi], True, True, n_total=n_c)
        Q, tape = self._queries(slot_inputs, True)
        E = E.contiguous()
        ptr_l, idx_l = restrict_csr(labels.ptr, labels.idx, lo, hi)
        y_base, y_pos = 0.0, 1.0
        if smoothing > 0:
            y_base, y_pos = (1.0 / n_c) * (1 - smoothing), (1.0 + 1.0 / n_c) * (1 - smoothing)
        q16, e16 = K.quantize(Q), K.quantize(E)
        loss_part, dS = K.score_bce(q16, e16, ptr_l, idx_l, y_base, y_pos)
        loss_sum = self.comm.all_reduce(loss_part.reshape(()).clone())
        g = 1.0 / float(normalizer_loss)
        dQ = K.gemm_nt(dS, K.ColMajor(e16), alpha=g)
        self.comm.all_reduce(dQ)
        dE = K.gemm_nt(dS.T, K.ColMajor(q16), alpha=g, splits=1)
        grads = {k: torch.zeros_like(self.p[k]) for k in self.trainable}
        # candidate side: local share -> all-reduce (batch-norm sums are reduced inside _encode_backward)
        self._encode_backward(dE, c_cand, grads)
        self.comm.all_reduce(grads["entity_embedding.weight"])
        # query side: identical on every rank, added once
        row = 0
        for fold, a, r, c_a, c_r in tape:
            b = a.size(0)
            ga, gr = K.fold_query_bwd(fold, a.contiguous(), r.contiguous(), dQ[row:row + b].contiguous())
            self._encode_backward(ga, c_a, grads)
            self._encode_backward(gr, c_r, grads)
            row += b
        for k in self.trainable:
            K.adagrad_dense(self.p[k], grads[k], self.state_sum[k], self.lr, self.eps, self.wd)
        return loss_sum

    # ---- filtered evaluation -----------------------------------------------------------------
    def candidate_block(self) -> torch.Tensor:
        """Eval-mode encode of this rank's block of ALL mentions (the sharded form of precompute_embeddings_from_tokens,
        openkge/model.py:670-712), cached until the next training step."""
        if self._eval_cache is None:
            lo, hi = shard_bounds(self.N, self.world, self.rank)
            ids = torch.arange(self.offset + lo, self.offset + hi, dtype=torch.int32,
                               device=self.p["entity_embedding.weight"].device)
            E = self._encode("entity", ids, False, False)[0].contiguous()
            self._eval_cache = (E, lo, hi, self.K.quantize(E, split=True) if hi > lo else None)
        return self._eval_cache

    def eval_counts(self, batch):
        slot_inputs, _, _, _, label_ids, filt, _ = batch
        E, lo, hi, e16 = self.candidate_block()
        Q, _ = self._queries(slot_inputs, False)
        return sharded_rank_counts(self.K, self.comm, Q.contiguous(), E, lo, hi, self.rank, label_ids, filt, e16=e16)

    def evaluate_batch(self, batch):
        _, greater, equal = self.eval_counts(batch)
        return metrics_from_counts(greater, equal)
