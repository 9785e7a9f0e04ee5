"""CUDA-graph replay of the training step.

``Trainer.compute_one_batch`` queues 20-40 kernels per step, three of them large; the rest are 2-10 us kernels whose
launch gaps add up (8 % of the step at the 1 M-entity config, most of it at FB15k-237 size). ``GraphedTrainStep`` captures
the whole step once — encode, fused scoring + loss, backward contractions, optimizer — and replays it per batch: the
host copies the batch into static device buffers and issues ONE graph launch.

What makes the step capturable: no shape in it depends on device data (CSR labels are read through a row pointer, the
positives buffer has a fixed capacity), nothing in it synchronises with the host (``sync_loss=False``), the kernels are
ordinary stream launches through the C ABI, and the two prefix kinds of a batch are folded from per-batch row kinds
that live in a static device tensor when the scorer treats them differently.

Scope: Lookup and token-pooling embedders (DistMult and ComplEx, with or without dropout and batch norm: C1-C4). Dropout
launches take their step index from a device counter that the graph increments, so every replay draws fresh masks;
batch-norm statistics are taken over row segments whose bounds (the po / sp split) are a static device tensor.
Batch-shared candidate lists (openkge/dataset.py:813-860) change length per batch: the list is padded to a fixed capacity
(``max_candidates``), the real count is device data that bounds the batch-norm statistics, masks the padded columns in the
loss epilogue (no loss term, zero gradient) and scales the seed gradient 1 / (B * N). Projections, the N3 hook, label
smoothing over batch-shared lists and gradient accumulation raise ``GraphCaptureUnsupported`` and the caller keeps using
``compute_one_batch``.
"""
from __future__ import annotations

from typing import Optional

import torch

from .dataset import AllEntityIds, CSRMatrix, PackedBatch


# Stream capture polices "potentially unsafe" CUDA runtime calls (allocations, event queries, ...) and invalidates the
# capture when one happens -- in the default "global" mode even when ANOTHER thread makes it. A training process has such
# threads by design: the loader's prefetch thread pins host memory while the training thread captures the step
# (train_epoch captures lazily from the first batch), the host allocator polls the events of recycled pinned blocks, and
# autograd runs the captured backward on its own worker thread (which rules out "thread_local"). "relaxed" lifts the
# policing; what the captured region itself does is under our control: kernel launches, NCCL collectives and allocations
# from the graph's private pool.
CAPTURE_MODE = "relaxed"


class GraphCaptureUnsupported(RuntimeError):
    pass


class GraphedTrainStep:
    def __init__(self, trainer, rows: int, max_positives: int, example_batch, max_candidates: Optional[int] = None,
                 preserve_state: bool = False):
        """``preserve_state``: the capture runs warm-up steps on ``example_batch``; True restores parameters, buffers
        (batch-norm statistics) and optimizer state afterwards, so that a training loop can create the graphed step lazily
        from its first batch without training on it four extra times."""
        model = trainer.model
        ds = trainer.train_dataset
        from .optim import Adagrad
        for regime in trainer.optimizers:
            opt = regime.optimizer
            if not isinstance(opt, Adagrad) or any(g.get("lr_decay", 0) != 0 for g in opt.param_groups):
                # the step number (Adam's bias correction, Adagrad's lr decay) is a host scalar baked into the launches
                raise GraphCaptureUnsupported("only Adagrad without lr decay is captured (step-dependent host scalars)")
        shard = getattr(model, "_shard", None)
        if shard is not None and shard.comm.on:
            import torch.distributed as dist
            if dist.get_backend(shard.comm.group) != "nccl":
                raise GraphCaptureUnsupported("the sharded step is captured with NCCL collectives only")
        if getattr(trainer, "data_parallel", False):
            import torch.distributed as dist
            if dist.get_backend() != "nccl":
                raise GraphCaptureUnsupported("the data-parallel step is captured with NCCL collectives only")
        self.shared = bool(ds.use_batch_shared_entities)
        if self.shared and trainer.model_with_loss.bce_label_smoothing > 0:
            raise GraphCaptureUnsupported("label smoothing over batch-shared candidates needs the count on the host")
        if self.shared and not isinstance(trainer.loss, torch.nn.BCEWithLogitsLoss):
            raise GraphCaptureUnsupported("batch-shared candidate lists are captured for the BCE loss only")
        if hasattr(model, "_lookup_batch"):
            if self.shared and getattr(model, "batch_norm", False):
                raise GraphCaptureUnsupported("Lookup + batch norm over a padded candidate list")
            if getattr(model, "project_entity", False) or getattr(model, "project_relation", False):
                raise GraphCaptureUnsupported("projections encode the po and sp blocks separately")
            if getattr(model, "normalize", "") == "norm" or getattr(model, "l2_reg", 0) > 0:
                raise GraphCaptureUnsupported("normalisation / N3 hook are not part of the captured step")
            self.has_batch_norm = bool(getattr(model, "batch_norm", False))
        elif hasattr(model, "_encode_rows"):
            if getattr(model, "relation_projection", None) is not None or getattr(model, "entity_projection", None) is not None:
                raise GraphCaptureUnsupported("projections encode the po and sp blocks separately")
            self.has_batch_norm = getattr(model, "normalize", None) == "batchnorm"
        else:
            raise GraphCaptureUnsupported("only Lookup and token-pooling embedders are captured")
        if trainer.batch_size_for_backward != ds.batch_size:
            raise GraphCaptureUnsupported("gradient accumulation")
        self.trainer, self.rows, self.capacity = trainer, int(rows), int(max_positives)
        dev = next(model.parameters()).device
        if self.shared:
            n_cols = int(max_candidates if max_candidates is not None else 2 * example_batch[6].numel())
            n_cols = (n_cols + 31) // 32 * 32
        else:
            n_cols = ds.index.n_cols
        self.n_cols = n_cols
        # the integer inputs of a step live in ONE device buffer [ent | rel | ptr | idx] (dataset.packed_layout): a batch that
        # carries the same layout on the host (dataset.PackedBatch, what the loaders yield) arrives with a single H2D copy
        from .dataset import packed_layout
        _, o_rel, o_ptr, o_idx = packed_layout(rows)
        self._o_idx = o_idx
        self.stage = torch.zeros(o_idx + self.capacity, dtype=torch.int32, device=dev)
        self.ent = self.stage[:rows].view(rows, 1)
        self.rel = self.stage[o_rel:o_rel + rows].view(rows, 1)
        self.ptr = self.stage[o_ptr:o_ptr + rows + 1]
        self.n_po_dev = self.stage[o_ptr + rows + 1:o_ptr + rows + 2]      # po rows of the loaded batch (they come first)
        self.idx = self.stage[o_idx:]
        self.idx.fill_(-1)
        labels = CSRMatrix(self.ptr, self.idx, (rows, n_cols))
        self.model = model
        self.asymmetric = model.fold_po != model.fold_sp
        self.kinds = torch.full((rows,), int(model.fold_sp), dtype=torch.int32, device=dev)
        self.dropout_step = torch.zeros((), dtype=torch.int64, device=dev)
        self.has_dropout = any(getattr(model, k, 0) for k in ("dropout", "input_dropout", "relation_dropout",
                                                              "relation_input_dropout", "entity_dropout"))
        # [begin, end) row ranges of the batch-norm statistics, refreshed by load(). Lookup models: the po and the sp block
        # of the query rows. Token models encode candidates and query rows in one call: candidates (the real ones of a
        # padded batch-shared list), po block, sp block behind them; then the two blocks of the relation rows.
        self.token_model = not hasattr(model, "_lookup_batch")
        n = n_cols
        self.segments = torch.tensor([0, n, n, n, n, n + rows, 0, 0, 0, rows] if self.token_model else [0, 0, 0, rows],
                                     dtype=torch.int32, device=dev)
        # all rows go in as one block; the prefix kind of every row (ComplEx folds po and sp rows differently) is data in
        # `kinds`, so the captured launches do not depend on the po / sp split of a batch
        if self.shared:
            # candidate ids padded with PAD (id 0); count, (0, count) and 1 / (rows * count) live next to them
            self.cand = torch.zeros((n_cols, 1), dtype=torch.int32, device=dev)
            self.cand_count = torch.zeros(1, dtype=torch.int32, device=dev)
            self.seed = torch.zeros((), dtype=torch.float32, device=dev)
            candidates = self.cand
        else:
            candidates = AllEntityIds(ds.index.offset, n_cols)
        self.static_batch = ([None, (self.ent, self.rel)], rows * n_cols, 0.0, labels, None, None, candidates)
        self.normalizer_loss = rows * n_cols
        self._hparams = self._hyper_parameters()
        self.sparse = bool(getattr(trainer, "sparse_exchange", False)) and self.shared
        if self.sparse:
            # data-parallel step with the touched-row exchange: the number of exchanged rows is data. A first graph marks and
            # numbers the rows (one small all-reduce) and copies their count to the host; the host picks the smallest
            # captured capacity that holds them and replays that variant of the step (captured on first use).
            self._init_sparse(example_batch, preserve_state)
            return
        snapshot = self._snapshot() if preserve_state else None
        self.load(example_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        def body():
            self._derive_split(self.cand_count if self.shared else None, self.n_po_dev)
            self._eager()

        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up: lazy initialisations, allocator, optimizer state
                body()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=CAPTURE_MODE):
            body()
            self.loss = trainer.last_loss            # device tensor owned by the graph's memory pool
        self._steps_per_replay = 1
        if snapshot is not None:
            self._restore(snapshot)

    # ---- data-parallel step with the touched-row exchange -------------------------------------------------------------------
    # capacity classes: 4 % of all token rows, then steps of 8 % up to all of them. Variants are captured on first use, so a
    # run holds the one or two classes its union size moves between (the size is stable to a few percent from step to step)
    _CAP_FRACTIONS = tuple(min(1.0, 0.04 * 1.08 ** i) for i in range(43))

    def _init_sparse(self, example_batch, preserve_state: bool) -> None:
        trainer = self.trainer
        total = trainer.union_state()["rows_total"]
        self._caps = sorted({min(total, (int(total * f) + 255) // 256 * 256) for f in self._CAP_FRACTIONS})
        self._variants, self.captured_capacities = {}, []
        dev = self.stage.device
        self._count_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._count_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._count_event = torch.cuda.Event()
        self.loss_per_label = torch.zeros((), dtype=torch.float32, device=dev)
        self.load(example_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._mark()
        torch.cuda.current_stream().wait_stream(side)
        self.mark_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.mark_graph, capture_error_mode=CAPTURE_MODE):
            self._mark()
        self._replay_sparse(self.mark_graph, restore=True)           # captures the variant the example batch needs
        self._steps_per_replay = 1

    def _mark(self) -> None:
        """Part of the marking graph: union numbering of the token rows this step touches, count to the host."""
        count = self.trainer.mark_union(torch.cat([self.cand.view(-1), self.ent.view(-1)]), self.rel.view(-1))
        self._count_dev.copy_(count.reshape(1))
        self._count_host.copy_(self._count_dev, non_blocking=True)

    def _variant(self, cap: int):
        """The step captured for ``cap`` exchanged rows (all ranks capture the same variant at the same step: the count is
        the same on every rank). State is snapshotted around the warm-up and capture."""
        if cap in self._variants:
            return self._variants[cap]
        trainer = self.trainer
        snapshot = self._snapshot()
        buf = trainer.union_buffer(trainer.union_state()["rows_total"] + 1)

        def body():
            buf[:cap].zero_()
            self._derive_split(self.cand_count, self.n_po_dev)
            self._eager()
            self.loss_per_label.copy_(self.trainer.last_loss.reshape(()) * self.seed)

        trainer.set_union(cap)
        trainer._external_union = True
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode=CAPTURE_MODE):
                body()
                loss = trainer.last_loss
        finally:
            trainer._external_union = False
            trainer.set_union(None)
        self._restore(snapshot)
        self._variants[cap] = (graph, loss)
        self.captured_capacities.append(cap)
        return self._variants[cap]

    def _replay_sparse(self, first_graph, restore: bool = False) -> None:
        """first_graph (marking, or device collate + marking) -> count on the host -> the variant that holds it."""
        first_graph.replay()
        self._count_event.record()
        self._count_event.synchronize()
        count = int(self._count_host[0])
        # the smallest variant already captured that holds the rows; a new one (an expensive capture in the middle of
        # training) only when none does, and then with 5 % headroom: the union size moves by a few percent from step to
        # step, and a step that exceeds its class by one row must not trigger a capture every other step
        held = [c for c in self._variants if c >= count]
        cap = min(held) if held else next(c for c in self._caps if c >= min(int(count * 1.05) + 1, self._caps[-1]))
        graph, self.loss = self._variant(cap)
        if not restore:
            graph.replay()
        self.last_union_rows = count

    def _derive_split(self, count_dev: Optional[torch.Tensor], n_po_dev: torch.Tensor) -> None:
        """Row kinds (ComplEx folds po and sp rows differently) and batch-norm segments from the number of po rows (and,
        for a padded batch-shared list, its real length), and the dropout step counter += 1: device data in, device data
        out, ONE launch (``okge_batch_layout``) at the head of the captured step."""
        if not (self.asymmetric or self.has_batch_norm or self.has_dropout):
            return
        from . import kernels as K
        K.batch_layout(n_po_dev, self.rows, self.n_cols, int(self.model.fold_po), int(self.model.fold_sp),
                       self.kinds if self.asymmetric else None, self.segments if self.has_batch_norm else None,
                       self.token_model, count_dev=count_dev, step_counter=self.dropout_step if self.has_dropout else None)

    def _hyper_parameters(self):
        return [(g["lr"], g["eps"], g["weight_decay"]) for r in self.trainer.optimizers for g in r.optimizer.param_groups]

    def _snapshot(self):
        model_state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        opt_state = []
        for regime in self.trainer.optimizers:
            for p, st in regime.optimizer.state.items():
                opt_state.append((st, {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}))
        return model_state, opt_state, (self.model._dropout_calls, self.trainer.training_steps, self.dropout_step.clone())

    def _restore(self, snapshot) -> None:
        model_state, opt_state, (calls, steps, dropout_step) = snapshot
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model_state[k])                      # in place: the graph holds these addresses
            for st, saved in opt_state:
                for k, v in saved.items():
                    if torch.is_tensor(v):
                        st[k].copy_(v)
                    else:
                        st[k] = v
        self.model._dropout_calls, self.trainer.training_steps = calls, steps
        self.dropout_step.copy_(dropout_step)        # (a variant captured in the middle of training keeps the stream position)
        from .functional import refresh_table_shadows
        refresh_table_shadows(self.model)            # fp16 table copies follow the restored weights (same buffers)
        torch.cuda.synchronize()

    def _eager(self):
        model = self.model
        saved = (model._graph_row_kinds, model._dropout_step_dev, model._dropout_calls, model._graph_segments,
                 model._graph_candidate_count)
        model._graph_row_kinds = self.kinds if self.asymmetric else None
        model._graph_segments = self.segments if self.has_batch_norm else None
        if self.shared:
            model._graph_candidate_count = self.cand_count
            self.trainer._graph_seed_gradient = self.seed
        if self.has_dropout:
            model._dropout_step_dev = self.dropout_step     # incremented by _derive_split: a new Philox position per replay
            model._dropout_calls = 0                 # the captured call indices restart every step
        try:
            self.trainer.compute_one_batch(self.static_batch, training=True, sync_loss=False)
        finally:
            (model._graph_row_kinds, model._dropout_step_dev, model._dropout_calls, model._graph_segments,
             model._graph_candidate_count) = saved
            self.trainer._graph_seed_gradient = None

    def load(self, batch) -> float:
        """Copies a collated (host or device) batch into the static buffers; returns its normalizer_metric."""
        if self._hyper_parameters() != self._hparams:
            raise RuntimeError("learning rate / eps / weight decay changed since the capture (they are baked into the "
                               "captured launches): create a new graphed step")
        if isinstance(batch, PackedBatch) and not self.shared:
            # fast path: one copy for all integer inputs (the 7-tuple's tensor views are never built)
            packed = batch.packed
            if packed.numel() - self._o_idx > self.capacity or batch.normalizer_loss != self.normalizer_loss \
                    or batch.rows != self.rows:
                raise ValueError(f"graphed step was captured for {self.rows} rows and {self.capacity} positives")
            self.stage[:packed.numel()].copy_(packed, non_blocking=True)
            self._set_split(batch.n_po)
            self._b_po = None                        # the device copy of n_po came with the packed buffer
            return batch.normalizer_metric
        slot_inputs, normalizer_loss, normalizer_metric, labels, _, _, shared_ids = batch
        po, sp = slot_inputs
        ent = [t for t in ((po[1] if po is not None else None), (sp[0] if sp is not None else None)) if t is not None]
        rel = [t for t in ((po[0] if po is not None else None), (sp[1] if sp is not None else None)) if t is not None]
        ent = ent[0] if len(ent) == 1 else torch.cat(ent)
        rel = rel[0] if len(rel) == 1 else torch.cat(rel)
        if ent.numel() != self.rows or (not self.shared and normalizer_loss != self.normalizer_loss):
            raise ValueError(f"graphed step was captured for {self.rows} rows")
        if self.shared:
            count = int(shared_ids.numel())
            if count > self.n_cols:
                raise ValueError(f"batch has {count} candidates, capacity is {self.n_cols}")
            if count != getattr(self, "_count", None):
                self.cand_count.fill_(count)                           # fill kernels: no host synchronisation
                self.seed.fill_(1.0 / float(self.rows * count))
                self._count = count
            self.cand[:count].copy_(shared_ids.reshape(-1, 1), non_blocking=True)
            # rows behind `count` keep stale ids of earlier batches: they are valid ids, encode to finite rows (zeros
            # after batch norm) and their columns are masked by the loss epilogue
            self.normalizer_loss = self.rows * count
        nnz = labels.idx.numel()
        if nnz > self.capacity:
            raise ValueError(f"batch has {nnz} positives, capacity is {self.capacity}")
        self.ent.copy_(ent.reshape(-1, 1), non_blocking=True)
        self.rel.copy_(rel.reshape(-1, 1), non_blocking=True)
        b_po = 0 if po is None else po[0].numel()
        self._set_split(b_po)
        if (self.asymmetric or self.has_batch_norm) and b_po != getattr(self, "_b_po", None):
            self.n_po_dev.fill_(b_po)
            self._b_po = b_po
        self.ptr.copy_(labels.ptr, non_blocking=True)
        self.idx[:nnz].copy_(labels.idx, non_blocking=True)
        return normalizer_metric

    def _set_split(self, b_po: int) -> None:
        """Rows [:b_po] of the loaded batch are po prefixes, the rest sp. The graph derives the row kinds / batch-norm
        segments from the device copy of b_po (``_derive_split``); a PackedBatch brings it along in its single copy."""
        if self.has_batch_norm and 1 in (b_po, self.rows - b_po):
            raise ValueError("Expected more than 1 value per channel when training (a one-row po or sp block)")

    def accepts(self, batch) -> bool:
        """True if ``batch`` fits the captured shapes (rows, positives / candidates within capacity, same optimizer
        hyper-parameters); a training loop runs the other batches (e.g. a ragged last one) through compute_one_batch."""
        if isinstance(batch, PackedBatch) and not self.shared:
            return (batch.rows == self.rows and batch.packed.numel() - self._o_idx <= self.capacity
                    and batch.normalizer_loss == self.normalizer_loss and self._hyper_parameters() == self._hparams)
        slot_inputs, normalizer_loss, _, labels, _, _, shared_ids = batch
        rows = sum(s[0].numel() for s in slot_inputs if s is not None)
        if rows != self.rows or labels.idx.numel() > self.capacity or self._hyper_parameters() != self._hparams:
            return False
        if self.shared:
            return shared_ids is not None and int(shared_ids.numel()) <= self.n_cols
        return normalizer_loss == self.rows * self.n_cols

    def __call__(self, batch) -> torch.Tensor:
        """One training step on ``batch``; returns the loss sum as a device tensor (valid until the next call)."""
        self.load(batch)
        # weights changed behind the graph since the last step (load_state_dict, a manual edit): the captured launches
        # read the tables' fp16 copies, which only the captured optimizer keeps current -- rebuild them, in place
        if getattr(self, "_shadowed", None) is None or self._replays_since_scan >= 64:
            # (parameters that carry an fp16 copy; re-scanned now and then: a shadow may be created lazily)
            self._shadowed = [p for p in self.model.parameters() if getattr(p, "_okge_shadow", None) is not None]
            self._replays_since_scan = 0
        self._replays_since_scan += 1
        for p in self._shadowed:
            sh = getattr(p, "_okge_shadow", None)
            if sh is not None and sh.op is not None and (sh.version != p._version or sh.ptr != p.data_ptr()):
                from .functional import table_operand
                sh.dirty = True
                table_operand(p, 0, split=sh.op.lo is not None)
        if self.sparse:
            self._replay_sparse(self.mark_graph)
        else:
            self.graph.replay()
        self._replays = getattr(self, "_replays", 0) + 1
        if self._replays % 512 == 0:                 # re-derive the power-of-two scales of the fp16 table copies
            from .functional import refresh_table_shadows
            refresh_table_shadows(self.model)
        for regime in self.trainer.optimizers:       # keep the python-side step counters of the optimizer in line
            for st in regime.optimizer.state.values():
                if "step" in st:
                    st["step"] += 1
        return self.loss

    @staticmethod
    def _rank() -> int:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    # ---- collate on the device (batch-shared candidate lists) --------------------------------------------------------
    def enable_device_collate(self, index, min_size_batch_labels: int) -> None:
        """Captures a second graph: ``dataset.DeviceSharedCollate`` of the row indices in ``self.rows_dev`` (candidate list,
        CSR labels, prefix ids, po / sp kinds, batch-norm segments, 1 / (B * count): all written into the static buffers
        of the step) followed by the training step itself. A batch is then B row indices already on the device; no host
        collate, no pinned staging, no H2D copy (``step_rows``)."""
        from .dataset import DeviceSharedCollate
        if not self.shared:
            raise GraphCaptureUnsupported("the device collate builds batch-shared candidate lists")
        dev = self.ent.device
        self.collate = DeviceSharedCollate(index, min_size_batch_labels, cap_nnz=self.capacity, cap_cols=self.n_cols, device=dev,
                                           seed=(int(torch.initial_seed()) + 7919 * self._rank()) & 0x7FFFFFFF,
                                           rows_per_batch=self.rows)
        self.rows_dev = torch.zeros(self.rows, dtype=torch.int64, device=dev)
        if not hasattr(self, "loss_per_label"):      # (the capacity variants of the sparse exchange already write it)
            self.loss_per_label = torch.zeros((), dtype=torch.float32, device=dev)
        fold_po, fold_sp = int(self.model.fold_po), int(self.model.fold_sp)
        # constants the captured launches read: they must outlive this call (the graph holds their addresses)
        self._collate_consts = (torch.tensor(self.rows, dtype=torch.int64, device=dev),
                                torch.tensor(self.n_cols, dtype=torch.int64, device=dev),
                                torch.zeros((), dtype=torch.int64, device=dev))
        rows_t, ncols_t, zero = self._collate_consts
        # the collate writes the static buffers of the step directly
        static = dict(ent=self.ent.view(-1), rel=self.rel.view(-1), ptr=self.ptr, idx=self.idx, cand=self.cand.view(-1),
                      count=self.cand_count, inv_norm=self.seed.view(1))

        def run():
            out = self.collate(self.rows_dev, out=static)
            if self.sparse:                          # the step itself is one of the capacity variants (``_replay_sparse``)
                self.n_po_dev.copy_(out["b_po"].reshape(1))
                self._mark()
                return
            self.n_po_dev.copy_(out["b_po"].reshape(1))
            self._derive_split(self.cand_count, self.n_po_dev)
            self._eager()
            self.loss_per_label.copy_(self.trainer.last_loss.reshape(()) * self.seed)

        snapshot = self._snapshot()
        self.rows_dev.copy_(torch.randint(0, len(index), (self.rows,), device=dev))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.row_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.row_graph, capture_error_mode=CAPTURE_MODE):
            run()
        self._restore(snapshot)
        self.collate.scalars.zero_()
        self._count = None                           # the static candidate buffers were overwritten: refresh on the next load()

    def step_rows(self, batch, sync_loss=True):
        """One training step on the prefix rows ``batch.rows`` (device int64 [B]), collated on the device inside the graph.
        Same return contract as ``step``; the loss meter gets loss / (B * N_candidates) of every step."""
        from .metrics import MetricResult
        if self._hyper_parameters() != self._hparams:
            raise RuntimeError("learning rate / eps / weight decay changed since the capture: create a new graphed step")
        self.rows_dev.copy_(batch.rows, non_blocking=True)
        if self.sparse:
            self._replay_sparse(self.row_graph)
        else:
            self.row_graph.replay()
        for regime in self.trainer.optimizers:
            for st in regime.optimizer.state.values():
                if "step" in st:
                    st["step"] += 1
        trainer, result = self.trainer, MetricResult()
        weight = self.rows * max(self.collate.min_size, 1)
        if sync_loss == "lagged":
            prev = trainer._read_lagged_loss()
            if prev is not None:
                result["loss"].update(*prev)
            host = trainer._pinned_scalar()
            host.copy_(self.loss_per_label, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            trainer._lagged_loss = (host, event, 1.0)
            trainer._lagged_weight = weight
        elif sync_loss:
            result["loss"].update(float(self.loss_per_label.item()), weight)
        return result, None

    def step(self, batch, sync_loss=True):
        """Same contract as ``Trainer.compute_one_batch(batch, training=True, sync_loss=...)``:
        returns (MetricResult with the loss meter, normalizer_metric)."""
        from .metrics import MetricResult
        trainer = self.trainer
        normalizer_metric = batch[2]
        loss = self(batch)
        result = MetricResult()
        if sync_loss == "lagged":
            prev = trainer._read_lagged_loss()
            if prev is not None:
                result["loss"].update(*prev)
            host = trainer._pinned_scalar()
            host.copy_(loss.reshape(()), non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            trainer._lagged_loss = (host, event, self.normalizer_loss)
        elif sync_loss:
            result["loss"].update(loss.item() / self.normalizer_loss, self.normalizer_loss)
        return result, normalizer_metric
