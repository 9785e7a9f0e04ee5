"""Loss module and batch / eval loop of the B200 path — drop-in for ``openkge/trainer.py``.

``AddLossModule`` keeps the reference's constructor and ``forward`` signature and its
``(loss_sum, hook_loss, all_outputs)`` return triple (openkge/trainer.py:37-56, 107-111) but runs the
po + sp prefix scoring and the loss as ONE fused tensor-core pass over the candidate table: the
``[B, N]`` score matrix of ``torch.cat(all_outputs)`` (:91) is never written, labels are CSR.
``Trainer.compute_one_batch`` follows openkge/trainer.py:181-272 line by line.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
from torch.nn import BCEWithLogitsLoss, KLDivLoss

from . import functional as Fn
from . import kernels as K
from .dataset import CSRMatrix, DeviceRows, PrefixScores
from .metrics import MetricResult
from .optim import OptimRegime


class AddLossModule(nn.Module):
    """Wraps a model to add the loss (openkge/trainer.py:32-113)."""

    def __init__(self, model, loss, bce_label_smoothing: float = 0.0, materialize_outputs: bool = False):
        super().__init__()
        self.model = model
        self.loss = loss
        self.bce_label_smoothing = bce_label_smoothing
        # True: all_outputs is the dense [B, N] matrix like the reference; False (default): None while
        # training (the reference never reads it there, :248-257) and a lazy PrefixScores in eval.
        self.materialize_outputs = materialize_outputs
        self.defer_eval_loss = False      # set by Trainer.compute_one_batch around its evaluation call

    def forward(self, inputs, labels, use_batch_shared_entities, batch_shared_entities, epoch=-1,
                input_style_triple_or_prefix="triple"):
        allowed = ["triple", "right_and_left_prefix"]
        if input_style_triple_or_prefix not in allowed:
            raise Exception("input_style_triple_or_prefix not in {}".format(allowed))
        if input_style_triple_or_prefix != "right_and_left_prefix":
            return None  # the reference implements only the prefix style (:58-113)
        if not (isinstance(self.loss, BCEWithLogitsLoss) or isinstance(self.loss, KLDivLoss)):
            raise NotImplementedError(f"{self.loss} not supported. Please choose either BCEWithLogitsLoss or KLDivLoss")
        if getattr(self.loss, "reduction", "sum") != "sum":
            raise NotImplementedError("the reference runs both losses with reduction='sum' (scripts/train.py:107-110)")

        model = self.model
        po_input, sp_input = inputs[0], inputs[1]
        # candidates: batch-shared ids, or every real entity (:75-82). In 1-vs-all mode the ids tensor is
        # arange(2, entities_size) (openkge/dataset.py:872) and is not needed on the device.
        candidate_ids = None
        if use_batch_shared_entities and batch_shared_entities is not None:
            candidate_ids = batch_shared_entities.reshape(-1)
        E, Q = model.encode_queries(po_input, sp_input, candidate_ids)   # rows ordered po first, then sp (:69-71)
        E = E.reshape(-1, E.size(-1))

        hook_loss = None
        if hasattr(model, "after_batch_loss_hook"):
            hook_loss = model.after_batch_loss_hook(epoch)

        if not isinstance(labels, CSRMatrix):
            labels = CSRMatrix.from_dense(labels.to(Q.device))   # reference-format dense [B, N] labels
        # entity table partitioned over ranks (model.shard_entities): E is this rank's block of the candidates. The labels
        # keep their row pointer, columns of other blocks become -1; the queries pass through ReplicatedInput (its
        # backward all-reduces dQ); the rank-local loss sums are all-reduced below.
        shard = getattr(model, "_shard", None) if candidate_ids is None else None
        N = E.size(0) if shard is None else labels.shape[1]          # all candidates of the job (label smoothing, :103-105)
        if shard is not None:
            from .sharded import restrict_csr
            if labels.shape != (Q.size(0), model.train_data.entities_size - model.train_data.min_entities_size):
                raise ValueError(f"labels {labels.shape} do not match the sharded candidate table")
            labels = CSRMatrix(*restrict_csr(labels.ptr, labels.idx, shard.lo, shard.hi), (Q.size(0), E.size(0)))
            Q = Fn.ReplicatedInput.apply(Q, shard.comm)
        if labels.shape != (Q.size(0), E.size(0)):
            raise ValueError(f"labels {labels.shape} do not match scores {(Q.size(0), E.size(0))}")
        pad = getattr(model, "grad_pad_rows", 0) if (candidate_ids is None and model.training) else 0
        # opt-in (Trainer args["fused_entity_update"]): leave dE = dS^T Q to the optimizer, which fuses it with its
        # Adagrad step; possible only when the candidate operand is the parameter table itself
        # (at every batch size: with fp16 operands the fused kernel also wins in the tensor-bound regime of large sharded
        # batches, see sharded.EntityShardedLookupModel)
        defer = bool(getattr(model, "fused_entity_update", False) and getattr(model, "_candidates_are_raw_table", False)
                     and torch.is_grad_enabled())
        # Evaluation inside Trainer.compute_one_batch (``defer_eval_loss``): the loss is produced by the ranking pass over the
        # same scores (okge_score_bce_rank), see dataset.PrefixScores.pending_loss; the returned loss tensor (float64,
        # 0-dim) is filled when ``compute_metrics`` / ``rank_answers`` / ``ensure_loss`` run on ``all_outputs``.
        deferred_eval = (self.defer_eval_loss and not model.training and not torch.is_grad_enabled() and shard is None
                         and isinstance(self.loss, BCEWithLogitsLoss) and not self.materialize_outputs)
        # fp16 operand of the candidates when the model already keeps one (the table's shadow copy, the eval cache);
        # evaluation ranks in split precision (hi + lo planes: fp32-grade scores) unless the model opts out
        split_eval = bool(getattr(model, "eval_split_precision", True)) and not model.training
        operand_of = getattr(model, "scoring_operand", None)
        e16 = operand_of(E, split=False) if operand_of is not None else None
        # input dropout of the candidate rows (the reference drops the whole candidate matrix of a step, model.py:461-470
        # via _get_all): the model left E as the raw table and named the dropout (model._candidate_dropout); it is applied
        # to the fp16 operand (one pass over the table's fp16 copy) and, in the backward, to the gradient tile
        cand_drop = getattr(model, "_candidate_dropout", None) if model.training else None
        if cand_drop is not None:
            if e16 is None:
                raise RuntimeError("candidate dropout was deferred to the scoring pass, but the table has no fp16 operand")
            e16 = K.mask_dropout_f16(e16, *cand_drop)
            e16.dropout = cand_drop
        e16_eval = None
        if not model.training and not self.materialize_outputs:
            e16_eval = operand_of(E, split=split_eval) if operand_of is not None else None
        if deferred_eval:
            y_base, y_pos = 0.0, 1.0
            if self.bce_label_smoothing > 0:
                y_base = (1.0 / N) * (1 - self.bce_label_smoothing)
                y_pos = (1.0 + 1.0 / N) * (1 - self.bce_label_smoothing)
            out = torch.zeros(1, dtype=torch.float64, device=Q.device)
            pending = dict(ptr=labels.ptr, idx=labels.idx, y_base=y_base, y_pos=y_pos, out=out)
            return out.reshape(()), hook_loss, PrefixScores(Q.detach(), E.detach(), pending_loss=pending, e16=e16_eval,
                                                            split=split_eval)
        if isinstance(self.loss, KLDivLoss):
            result = Fn.ScoreKLLoss.apply(Q, E, labels.ptr, labels.idx, pad, defer, e16, shard)
        else:
            y_base, y_pos = 0.0, 1.0
            if self.bce_label_smoothing > 0:                     # y <- (y + 1/N)(1 - eps), :103-105
                y_base = (1.0 / N) * (1 - self.bce_label_smoothing)
                y_pos = (1.0 + 1.0 / N) * (1 - self.bce_label_smoothing)
            # CUDA-graph replay of batch-shared candidate lists: E is padded to a fixed capacity, the real count is data
            n_cols_dev = getattr(model, "_graph_candidate_count", None) if candidate_ids is not None else None
            if n_cols_dev is not None and self.bce_label_smoothing > 0:
                raise NotImplementedError("label smoothing needs the candidate count on the host")
            result = Fn.ScoreBCELoss.apply(Q, E, labels.ptr, labels.idx, y_base, y_pos, pad, defer, n_cols_dev, e16)

        if shard is not None:
            result = Fn.AllReduceSum.apply(result, shard.comm)
            if self.materialize_outputs:
                raise NotImplementedError("a dense [B, N] score matrix does not exist when the candidates are sharded")
        if self.materialize_outputs:
            all_outputs = Fn.ScoreMatrix.apply(Q, E)
        elif model.training:
            all_outputs = None
        else:
            all_outputs = PrefixScores(Q.detach(), E.detach(), e16=e16_eval, split=split_eval, shard=shard)
        return result, hook_loss, all_outputs


class Trainer(object):
    """Batch / eval loop (openkge/trainer.py:115-369). Early stopping, checkpoint rotation and the results
    CSV of the reference (:371-638) are bookkeeping outside the accelerated path and are not rebuilt."""

    def __init__(self, args, model, loss, train_dataset, validation_dataset, train_loader=None):
        self.args = args
        self.train_dataset = train_dataset
        self.validation_dataset = validation_dataset
        # A torch.distributed job with more than one rank (scripts/train.py:86-124 would wrap the model in DataParallel
        # here): 1-vs-all Lookup models partition their entity table over the ranks instead (model.shard_entities; every
        # rank is fed the same batches). args["entity_sharding"] = False keeps the table replicated.
        import torch.distributed as dist
        if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and args.get("entity_sharding", True)
                and hasattr(model, "shard_entities") and getattr(model, "_shard", None) is None
                and not getattr(train_dataset, "use_batch_shared_entities", False)):
            model.shard_entities(dist.get_rank(), dist.get_world_size())
        # Batch-shared candidate lists (the OLPBench configurations) in a multi-rank job: plain data parallelism. Every rank
        # draws its own batches (a loader seeded per rank) and candidate lists, the gradients are averaged over the ranks
        # before the optimizer step (what DataParallel / DistributedDataParallel do for scripts/train.py:118-124);
        # args["data_parallel"] = False turns it off.
        self.data_parallel = bool(
            dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and args.get("data_parallel", True)
            and getattr(train_dataset, "use_batch_shared_entities", False))
        self.optimizers: List[OptimRegime] = OptimRegime.setup_optimizer_regime(args=args, model=model)
        self.model = model
        self.loss = loss
        self.model_with_loss = AddLossModule(self.model, self.loss, args.get("bce_label_smoothing", 0.0))
        self.training_steps = 0
        self.len_train_batches = 1 if train_loader is None else len(train_loader)
        bsb = getattr(train_dataset, "batch_size_for_backward", None)
        self.batch_size_for_backward = bsb if bsb is not None else train_dataset.batch_size
        self.batch_size_for_backward_accumulated = 0
        # "fused_entity_update" (not a reference option): the entity table's 1-vs-all gradient is applied by the
        # optimizer straight from the dE contraction (okge_gemm_adagrad) and `.grad` of that table stays None. It needs
        # one backward per step and no gradient clipping (both read the dense gradient), else it is ignored.
        clip = args.get("grad_clip")
        fused = bool(args.get("fused_entity_update", False))
        if fused and ((clip is not None and clip > 0) or self.batch_size_for_backward != train_dataset.batch_size):
            fused = False
        self.model.fused_entity_update = fused
        # token models: the same option keeps the token tables' gradients compact (functional.SlotTableGrad) whenever a
        # step touches fewer token slots than the table has rows; needs our Adagrad (it consumes the deferred gradient)
        from .optim import Adagrad as _Adagrad
        slot_ok = fused and all(isinstance(r.optimizer, _Adagrad) or r.optimizer.__class__.__name__ == "Adam"
                                for r in self.optimizers)
        # Data-parallel token models exchange only the token rows some rank touched in the step (functional.UnionSlots)
        # instead of all-reducing the dense [V, D] gradients: 107 / 177 / 269 MB instead of 512 MB per step at 2 / 4 / 8
        # ranks on the OLPBench-shaped workload. Needs what the compact gradient needs (our optimizers, no clipping, no
        # accumulation) and one row width for both tables (they share the exchange buffer);
        # args["sparse_gradient_exchange"] = False keeps the dense all-reduce.
        self.sparse_exchange = bool(
            self.data_parallel and hasattr(model, "_encode_rows") and slot_ok and args.get("sparse_gradient_exchange", True)
            and model.entity_embedding.weight.size(1) == model.relation_embedding.weight.size(1))
        self._external_union = False
        if hasattr(model, "_encode_rows"):
            for emb in (model.entity_embedding, model.relation_embedding):
                # data parallel: the exchange needs the union numbering (or, without it, the dense [V, D] gradients)
                emb.weight._okge_slot_update = bool(slot_ok) and not self.data_parallel

    @property
    def epoch(self):
        return math.floor(self.training_steps / (self.len_train_batches + 1)) + 1

    def _seed_gradient(self, normalizer_loss, like: torch.Tensor) -> torch.Tensor:
        """Device scalar 1 / normalizer_loss, cached per value (1-vs-all batches all share B * N)."""
        cache = self.__dict__.setdefault("_seed_gradients", {})
        key = (float(normalizer_loss), like.dtype, like.device)
        if key not in cache:
            if len(cache) > 256:
                cache.clear()
            cache[key] = torch.full((), 1.0 / float(normalizer_loss), dtype=like.dtype, device=like.device)
        return cache[key]

    def _clip_grad_norm(self, max_norm: float) -> None:
        """torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) (openkge/trainer.py:240-242); with a sharded entity
        table the squared norm of its local gradient blocks is summed over the ranks first, so every rank scales by the
        norm of the whole job's gradient."""
        shard = getattr(self.model, "_shard", None)
        if shard is None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm)
            return
        table = self.model.entity_embedding.weight
        grads = [p.grad for p in self.model.parameters() if p.grad is not None and p is not table]
        sq = torch.stack([g.detach().float().pow(2).sum() for g in grads]).sum() if grads else \
            torch.zeros((), device=table.device)
        if table.grad is not None:
            ms = self.model.train_data.min_entities_size
            local = table.grad.detach()[ms:].float().pow(2).sum()
            if shard.rank == 0:
                local = local + table.grad.detach()[:ms].float().pow(2).sum()
            sq = sq + shard.comm.all_reduce(local.clone())
            grads = grads + [table.grad]
        coef = torch.clamp(max_norm / (sq.sqrt() + 1e-6), max=1.0)
        for g in grads:
            g.detach().mul_(coef)

    # ---- touched-row exchange of the token-table gradients (data-parallel token models) ---------------------------------
    def union_state(self) -> dict:
        """Static buffers of the union numbering: touch flags of both token tables in ONE int32 vector (one all-reduce), the
        slot maps, and the exchange buffer [rows of both tables + 1, D] (its first ``cap`` rows are used by a step)."""
        st = self.__dict__.get("_union")
        if st is None:
            m = self.model
            tables = [(m.entity_embedding.weight, m._entity_token_ids_i32), (m.relation_embedding.weight, m._relation_token_ids_i32)]
            dev = tables[0][0].device
            sizes = [int(w.size(0)) for w, _ in tables]
            flat = torch.zeros(sum(sizes), dtype=torch.int32, device=dev)
            flags = list(flat.split(sizes))
            st = self.__dict__["_union"] = dict(
                tables=tables, flat=flat, flags=flags, rows_total=sum(sizes),
                write=[torch.full((n,), -1, dtype=torch.int32, device=dev) for n in sizes],
                read=[torch.full((n,), -1, dtype=torch.int32, device=dev) for n in sizes],
                zero=torch.zeros((), dtype=torch.int64, device=dev), buf=None)
        return st

    def union_buffer(self, rows: int) -> torch.Tensor:
        st = self.union_state()
        if st["buf"] is None or st["buf"].size(0) < rows:
            w = st["tables"][0][0]
            st["buf"] = torch.zeros((rows, w.size(1)), dtype=torch.float32, device=w.device)
        return st["buf"]

    def mark_union(self, entity_ids: torch.Tensor, relation_ids: torch.Tensor) -> torch.Tensor:
        """Touch flags of this rank's step (the token rows of the given entity / relation ids), all-reduced (max) over the
        ranks, numbered: fills the slot maps for an unbounded capacity and returns the number of touched rows (0-dim int64
        device tensor). No host synchronisation: a CUDA graph can hold it."""
        import torch.distributed as dist
        from . import functional as Fn
        st = self.union_state()
        st["flat"].zero_()
        for flags, (_, id_rows), ids in zip(st["flags"], st["tables"], (entity_ids, relation_ids)):
            Fn.touch_flags(flags, id_rows, ids)
        dist.all_reduce(st["flat"], op=dist.ReduceOp.MAX)
        base = st["zero"]
        for flags, wm, rm in zip(st["flags"], st["write"], st["read"]):
            base = Fn.union_slots(flags, base, st["rows_total"], wm, rm)
        return base

    def set_union(self, cap: Optional[int]) -> None:
        """Arms (``cap`` rows of the exchange buffer) or disarms (None) the touched-row path of the token tables' backward."""
        from . import functional as Fn
        st = self.union_state()
        for (w, _), wm, rm in zip(st["tables"], st["write"], st["read"]):
            w._okge_union = None if cap is None else Fn.UnionSlots(wm, rm, self.union_buffer(st["rows_total"] + 1), cap)

    def _average_gradients(self) -> None:
        """Data-parallel step: mean of every parameter gradient over the ranks (NCCL all-reduce, AVG). The small tensors
        (batch-norm parameters, LSTM weights) travel in one flat buffer; the token tables go as they are, or -- with the
        touched-row exchange -- as the first ``cap`` rows of the exchange buffer."""
        import torch.distributed as dist
        from . import functional as Fn
        small, big = [], []
        union = None
        for p in self.model.parameters():
            d = getattr(p, "_okge_deferred", None)
            if isinstance(d, Fn.UnionSlotGrad) and p.grad is None:
                union = d.union
                continue
            if p.grad is None:
                # a parameter without gradient on this rank still has to take part (another rank may have one)
                p.grad = torch.zeros_like(p.data)
            (big if p.grad.numel() >= (1 << 20) else small).append(p.grad)
        op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        scale = 1.0 if op == dist.ReduceOp.AVG else 1.0 / dist.get_world_size()
        if union is not None:
            big.append(union.buf[:union.cap])
        for g in big:
            dist.all_reduce(g, op=op)
            if scale != 1.0:
                g.mul_(scale)
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            dist.all_reduce(flat, op=op)
            if scale != 1.0:
                flat.mul_(scale)
            off = 0
            for g in small:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def sync_replicas(self) -> None:
        """Sharded entity table: everything else (relation table, batch-norm parameters, their optimizer state) is
        replicated and updated from identical gradients, but float atomics (the scatter-add of repeated relation rows) sum
        in a run-dependent order, so the replicas drift apart by rounding noise. Broadcasting rank 0's copy now and then
        (``train_epoch``: every args["replica_sync_steps"] = 1024 steps) keeps them identical."""
        shard = getattr(self.model, "_shard", None)
        import torch.distributed as dist
        if self.data_parallel:                 # replicas of everything, batch-norm running statistics included
            for t in list(self.model.parameters()) + [b for b in self.model.buffers() if b.is_floating_point()]:
                dist.broadcast(t.data, src=0)
            for regime in self.optimizers:
                for st in regime.optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v) and v.is_floating_point() and v.is_cuda:
                            dist.broadcast(v, src=0)
            return
        if shard is None or not shard.comm.on:
            return
        table = self.model.entity_embedding.weight
        tensors = [p.data for p in self.model.parameters() if p is not table]
        tensors += [b for b in self.model.buffers() if b.is_floating_point()]
        for regime in self.optimizers:
            for p, st in regime.optimizer.state.items():
                if p is not table:
                    tensors += [v for v in st.values() if torch.is_tensor(v) and v.is_floating_point() and v.is_cuda]
        for t in tensors:
            dist.broadcast(t, src=0, group=shard.comm.group)

    def _pinned_scalar(self) -> torch.Tensor:
        """A pinned fp32 scalar for the asynchronous loss read of this step, from a small ring (a lagged loss is read one
        step later, so four slots never alias); allocating pinned memory per step costs more than the copy."""
        ring = self.__dict__.get("_pinned_ring")
        if ring is None:
            ring = self.__dict__["_pinned_ring"] = [torch.empty((), dtype=torch.float32, pin_memory=True) for _ in range(4)]
            self._pinned_next = 0
        self._pinned_next = (self._pinned_next + 1) % len(ring)
        return ring[self._pinned_next]

    def _read_lagged_loss(self):
        """Host value of the previous step's loss (its D2H copy was queued behind that step)."""
        pend, self._lagged_loss = getattr(self, "_lagged_loss", None), None
        if pend is None:
            return None
        host, event, normalizer = pend
        event.synchronize()
        weight = getattr(self, "_lagged_weight", None)
        self._lagged_weight = None
        return float(host.item()) / normalizer, (weight if weight is not None else normalizer)

    def flush_loss(self, metric_result: Optional[MetricResult] = None) -> Optional[MetricResult]:
        """Drains the loss of the last ``sync_loss="lagged"`` step into ``metric_result`` (a new one if None)."""
        metric_result = MetricResult() if metric_result is None else metric_result
        last = self._read_lagged_loss()
        if last is not None:
            metric_result["loss"].update(*last)
        return metric_result

    def compute_one_batch(self, data, training=True, sync_loss=True):
        """``sync_loss`` selects how the training loss reaches the host meter. True: ``loss.item()`` every step like
        the reference (:250), which drains the GPU before the next step can be queued. ``"lagged"``: every step's
        loss is still copied to the host (pinned buffer, asynchronous) but read one step later, so the host queues
        step i+1 while step i runs; ``flush_loss()`` returns the last one. False: no host read, empty loss meter.
        In evaluation ``"lagged"`` returns a closure instead of the MetricResult (see ``evaluate``)."""
        data_set = self.train_dataset if training else self.validation_dataset
        inputs, normalizer_loss, normalizer_metric, labels, label_ids, filter_mask, batch_shared_entities = \
            data_set.input_and_labels_to_device(data, training=training, device=data_set.device)

        union_armed = False
        if training and self.sparse_exchange and not self._external_union and isinstance(batch_shared_entities, torch.Tensor):
            # eager data-parallel step: number the touched token rows now; the exact count sizes the exchange (one host
            # read per step -- the graphed step keeps the count on the device and picks a captured capacity instead)
            po, sp = inputs
            ent = [batch_shared_entities.reshape(-1)] + [t.reshape(-1) for t in ((po[1] if po is not None else None),
                                                                                 (sp[0] if sp is not None else None)) if t is not None]
            rel = [t.reshape(-1) for t in ((po[0] if po is not None else None), (sp[1] if sp is not None else None)) if t is not None]
            count = int(self.mark_union(torch.cat(ent), torch.cat(rel)).item())
            cap = max(256, (count + 255) // 256 * 256)
            self.set_union(cap)
            self.union_buffer(cap)[:cap].zero_()
            union_armed = True

        self.model_with_loss.defer_eval_loss = not training       # eval: loss and ranking share one pass over the candidates
        try:
            loss, hook_loss, predictions = self.model_with_loss(
                inputs=inputs, labels=labels, batch_shared_entities=batch_shared_entities,
                use_batch_shared_entities=data_set.use_batch_shared_entities, epoch=self.epoch,
                input_style_triple_or_prefix=data_set.input_style)
        finally:
            self.model_with_loss.defer_eval_loss = False
        batch_size = len(labels)
        self.last_loss = None if loss is None else loss.detach()

        backward_loss, backward_scale = None, None
        if loss is not None:
            if training and hook_loss is None and loss.dim() == 0:
                # loss / normalizer_loss (:217-221) without sum / div nodes: the scale goes in as the seed gradient
                backward_loss, backward_scale = loss, self._seed_gradient(normalizer_loss, loss)
                if getattr(self, "_graph_seed_gradient", None) is not None:     # graphed.GraphedTrainStep: 1 / (B * N) is data
                    backward_scale = self._graph_seed_gradient
            else:
                backward_loss = loss.sum()
                if hook_loss is not None:
                    backward_loss = backward_loss + hook_loss
                backward_loss = backward_loss / normalizer_loss                   # :217-221

        if training:
            if backward_loss is None:
                return None, normalizer_metric
            if self.batch_size_for_backward_accumulated == 0:
                for optimizer in self.optimizers:
                    optimizer.zero_grad()
            backward_loss.backward(gradient=backward_scale)
            self.batch_size_for_backward_accumulated += batch_size
            if self.batch_size_for_backward_accumulated == self.batch_size_for_backward:
                if self.data_parallel:
                    self._average_gradients()
                for optimizer in self.optimizers:
                    clip = self.args.get("grad_clip")
                    if clip is not None and clip > 0:
                        self._clip_grad_norm(clip)
                    optimizer.step()
                    optimizer.zero_grad()
                if union_armed:
                    self.set_union(None)
                self.batch_size_for_backward_accumulated = 0
                metric_result = MetricResult()
                if sync_loss == "lagged":
                    prev = self._read_lagged_loss()
                    if prev is not None:
                        metric_result["loss"].update(*prev)
                    host = self._pinned_scalar()
                    host.copy_(loss.detach().reshape(()), non_blocking=True)
                    event = torch.cuda.Event()
                    event.record()
                    self._lagged_loss = (host, event, normalizer_loss)
                elif sync_loss:
                    # one host read per step, like the reference's loss.detach().item() (:250)
                    metric_result["loss"].update(loss.detach().item() / normalizer_loss, normalizer_loss)
                return metric_result, normalizer_metric
            return None, normalizer_metric

        if sync_loss == "lagged":
            # evaluation pipelined by one batch (Trainer.evaluate): queue the rank passes and an asynchronous copy of the
            # seven result scalars, hand back a closure that builds the MetricResult once the copy has landed
            from .dataset import metric_sums, metrics_from_sums, rank_answers
            _, greater, equal, _ = rank_answers(filter_mask, label_ids, predictions)
            if hasattr(predictions, "ensure_loss"):
                predictions.ensure_loss()
            n_q = int(greater.numel())
            vals = torch.cat([metric_sums(greater, equal) if n_q else torch.zeros(6, dtype=torch.float64, device=greater.device),
                              (loss.detach().double().reshape(1) if loss is not None else
                               torch.zeros(1, dtype=torch.float64, device=greater.device))])
            host = torch.empty(7, dtype=torch.float64, pin_memory=True)
            host.copy_(vals, non_blocking=True)
            event = torch.cuda.Event()
            event.record()

            def finish():
                event.synchronize()
                v = host.tolist()
                result = metrics_from_sums(v[:6], n_q)
                result["loss"].update(v[6] / normalizer_loss if loss is not None else 0, normalizer_loss)
                return result
            return finish, normalizer_metric
        metric_result = data_set.compute_metrics(filter_mask, label_ids, predictions)  # :263-267
        if hasattr(predictions, "ensure_loss"):
            predictions.ensure_loss()
        metric_result["loss"].update(loss.detach().item() / normalizer_loss if loss is not None else 0, normalizer_loss)
        return metric_result, normalizer_metric

    def make_graphed_step(self, example_batch, max_positives: Optional[int] = None, max_candidates: Optional[int] = None,
                          preserve_state: bool = False):
        """A CUDA-graph replay of ``compute_one_batch(training=True)`` for batches shaped like ``example_batch`` (see
        ``graphed.GraphedTrainStep``), or None when this model / dataset configuration cannot be captured."""
        from .graphed import GraphCaptureUnsupported, GraphedTrainStep
        labels = example_batch[3]
        rows = len(labels)
        cap = max_positives if max_positives is not None else max(4096, 4 * int(labels.idx.numel()))
        try:
            return GraphedTrainStep(self, rows, cap, example_batch, max_candidates, preserve_state=preserve_state)
        except GraphCaptureUnsupported:
            return None

    def train_epoch(self, data_loader, max_steps: Optional[int] = None) -> MetricResult:
        """compute_one_epoch(training=True) without the logging / periodic-eval generator (:274-361).

        With ``args["cuda_graph"] = True`` (not a reference option) the step is captured from the first full batch
        (model and optimizer state restored afterwards) and every batch that fits the captured shapes is one CUDA-graph
        launch (``graphed.GraphedTrainStep``); the others - a ragged last batch, or all of them when the configuration
        cannot be captured - go through ``compute_one_batch``. The loss meter then reads every loss one step late."""
        self.model_with_loss.train()
        total = MetricResult()
        use_graph = bool(self.args.get("cuda_graph", False))
        for step, batch in enumerate(data_loader):
            if max_steps is not None and step >= max_steps:
                break
            # the reference counts the step and refreshes the epoch length BEFORE it updates the regime (:295-299)
            self.training_steps += 1
            if hasattr(data_loader, "__len__"):
                self.len_train_batches = len(data_loader)
            for optimizer in self.optimizers:
                optimizer.update(self.epoch, self.training_steps)
            if self.training_steps % int(self.args.get("replica_sync_steps", 1024)) == 0:
                self.sync_replicas()
            if isinstance(batch, DeviceRows):            # collate on the device, inside the CUDA graph of the step
                result, _ = self._graphed_step_for_rows(batch).step_rows(batch, sync_loss="lagged")
                if result is not None:
                    total = total + result
                use_graph = True
                continue
            graphed = self._graphed_step_for(batch) if use_graph else None
            if graphed is not None:
                result, _ = graphed.step(batch, sync_loss="lagged")
            else:
                if use_graph:
                    total = self.flush_loss(total)          # keep the meter in step order around an eager batch
                result, _ = self.compute_one_batch(batch, training=True)
            if result is not None:
                total = total + result
        return self.flush_loss(total) if use_graph else total

    def _graphed_step_for_rows(self, batch):
        """The graphed step with the device collate enabled (``graphed.GraphedTrainStep.enable_device_collate``), created
        from the first batch: its rows are collated once on the host to size the static buffers."""
        g = getattr(self, "_graphed_step", None)
        if g is not None and g._hyper_parameters() != g._hparams:
            g = self._graphed_step = None
        if g is None or not hasattr(g, "row_graph"):
            ds = self.train_dataset
            if not ds.use_batch_shared_entities:
                raise NotImplementedError("device-side collate exists for batch-shared candidate lists")
            if g is None:
                example = ds.collate(batch.rows.cpu().numpy())
                # candidate capacity of the static buffers: args["device_collate_capacity"] (default 2.0) times the first
                # batch's list; batches with longer lists are cut to it and counted (DeviceSharedCollate.overflow)
                factor = float(self.args.get("device_collate_capacity", 2.0))
                cap = (int(factor * max(int(example[6].numel()), int(ds.min_size_batch_labels))) + 255) // 256 * 256
                g = self.make_graphed_step(example, max_positives=max(4096, 4 * int(example[3].idx.numel())), max_candidates=cap,
                                           preserve_state=True)
                if g is None:
                    raise NotImplementedError("this model / dataset configuration cannot be captured as a CUDA graph")
                self._graphed_step = g
            g.enable_device_collate(ds.index, ds.min_size_batch_labels)
        return g

    def _graphed_step_for(self, batch):
        """The cached graphed step if ``batch`` fits it; created (once) from the first batch that has the dataset's batch
        size. None: run the batch eagerly."""
        g = getattr(self, "_graphed_step", None)
        if g is None and not getattr(self, "_graph_unsupported", False):
            rows = sum(s[0].numel() for s in batch[0] if s is not None)
            if rows != self.train_dataset.batch_size:
                return None
            shared = batch[6]
            cand = None
            if self.train_dataset.use_batch_shared_entities and shared is not None:
                cand = 2 * int(shared.numel())
            g = self.make_graphed_step(batch, max_candidates=cand, preserve_state=True)
            if g is None:
                self._graph_unsupported = True
                return None
            self._graphed_step = g
        if g is None:
            return None
        if g._hyper_parameters() != g._hparams:             # a new phase of the optimizer regime: capture again
            self._graphed_step = None
            return self._graphed_step_for(batch)
        return g if g.accepts(batch) else None

    def evaluate(self, data_loader) -> MetricResult:
        """openkge/trainer.py:363-369."""
        self.model_with_loss.eval()
        total = MetricResult()
        pending = None
        with torch.no_grad():
            for batch in data_loader:
                # batch i+1 is queued before the results of batch i are read: the GPU never waits for the host
                finish, _ = self.compute_one_batch(batch, training=False, sync_loss="lagged")
                if pending is not None:
                    total = total + pending()
                pending = finish
            if pending is not None:
                total = total + pending()
        return total
