"""Scorer x embedder mixins of the B200 path — drop-in for ``openkge/model.py``.

Same plug-in mechanism as the reference: a model is ``class M(SomeScorer, SomeEmbedder)`` composed by
multiple inheritance (openkge/model.py:1006-1019), registered on ``Models`` (:1052-1066) and built as
``getattr(Models, name)(**model_config, train_data=meta)`` (scripts/train.py:86-88). Class names,
constructor keyword arguments, method names (``sp_prefix_score``, ``po_prefix_score``, ``get_all_obj``,
``precompute_batch_shared_inputs``, ``encode_subj/rel/obj``, ``_score``, ``after_batch_loss_hook``) and
state-dict keys (``entity_embedding.weight``, ``relation_embedding.weight``, ``entity_batchnorm.*``,
``entity_token_ids`` ...) are the reference's, so checkpoints and configs carry over.

What differs is what runs underneath: every lookup, pooling, fold, scoring contraction and its
backward is one of the sm_100a kernels of ``include/okge_b200.h`` (via ``functional``); there is no
CPU path. Batch norm / optional projections / L2-normalise stay PyTorch modules.

Additions to the reference API, used by the fused loss module (``trainer.AddLossModule``):
``encode_prefix_batch`` (all encodes of one batch from a single autograd node per table, in the
reference's call order) and ``sp_prefix_query`` / ``po_prefix_query`` (the folded query vectors).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import functional as Fn
from .dataset import PAD, EntityRelationDatasetMeta
from .kernels import FOLD_COMPLEX_PO, FOLD_COMPLEX_SP, FOLD_DISTMULT


def _flat2d(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(-1, t.size(-1))


class _Sequential(torch.nn.Sequential):
    """Sequential that skips ``None`` entries (utils/torch_nn_modules.py:1-20)."""

    def __init__(self, *mods):
        super().__init__(*[m for m in mods if m is not None])


class RelationModel(torch.nn.Module):
    """Base class (openkge/model.py:14-29)."""

    is_cuda = False
    rel_obj_cache = None
    subj_rel_cache = None
    # Evaluation contracts the scores in split precision (fp16 hi + lo planes, three tensor-core passes): fp32-grade
    # scores, so the filtered ranks follow the reference's fp32 scorer. False: single fp16 pass (TF32-grade, 3x faster).
    eval_split_precision = True

    def scoring_operand(self, E: torch.Tensor, split: bool = False):
        """The fp16 operand of the candidate matrix ``E`` if the model already keeps one (else None: the loss module
        quantizes ``E`` itself)."""
        return None

    def cuda(self, device=None):
        super().cuda(device=device)
        self.is_cuda = True
        return self

    def cpu(self):
        raise RuntimeError("the B200 hot path has no CPU execution mode")

    # dropout stream bookkeeping shared by the embedders: Philox (seed, offset) per call
    _dropout_seed: Optional[int] = None
    _dropout_calls = 0
    # CUDA-graph replay (graphed.GraphedTrainStep): the step index lives in device memory so that the replayed launches
    # draw fresh masks; the per-step call counter restarts with every captured step
    _dropout_step_dev: Optional[torch.Tensor] = None
    _graph_row_kinds: Optional[torch.Tensor] = None
    _graph_segments: Optional[torch.Tensor] = None
    # batch-shared candidate list padded to a fixed capacity (graph replay): the real count, device int32
    _graph_candidate_count: Optional[torch.Tensor] = None

    def _row_segments(self, bounds, device) -> torch.Tensor:
        """int32 device tensor of [begin, end) row ranges (flattened pairs): the segments of the batch-norm statistics.
        The reference normalises the candidates, the po block and the sp block of a batch in separate calls
        (openkge/trainer.py:69-87); one call over all rows with these ranges does the same arithmetic."""
        if self._graph_segments is not None:
            return self._graph_segments                      # static tensor refreshed by GraphedTrainStep.load
        cache = self.__dict__.setdefault("_segment_cache", {})
        key = (tuple(int(b) for b in bounds), str(device))
        if key not in cache:
            if len(cache) > 1024:
                cache.clear()
            cache[key] = torch.tensor(key[0], dtype=torch.int32, device=device)
        return cache[key]

    def _dropout_spec(self, p: float):
        """(p, seed, offset, step_dev) of the next dropout call: what ``_dropout`` hands to the kernel, also used by the
        encoders that fuse the dropout into the preceding batch-norm pass."""
        if self._dropout_seed is None:
            self._dropout_seed = int(torch.initial_seed()) & (2**63 - 1)
        self._dropout_calls += 1
        if self._dropout_step_dev is not None:
            return float(p), self._dropout_seed, (self._dropout_calls & 31) << 38, self._dropout_step_dev
        return float(p), self._dropout_seed, self._dropout_calls << 38, None

    def _dropout(self, x: torch.Tensor, p: float) -> torch.Tensor:
        if p <= 0 or not self.training:
            return x
        return Fn.Dropout.apply(x, *self._dropout_spec(p))


# ---------------------------------------------------------------------------------------------
# scorers
# ---------------------------------------------------------------------------------------------

class RelationScorer(RelationModel):
    """openkge/model.py:31-77."""

    fold_sp = FOLD_DISTMULT
    fold_po = FOLD_DISTMULT

    def forward(self, subj, rel, obj, **kwargs):
        return self.triple_score(self.encode_subj(subj), self.encode_rel(rel), self.encode_obj(obj), **kwargs)

    def triple_score(self, subj, rel, obj, **kwargs):
        return self._score(subj, rel, obj)

    def sp_prefix_score(self, subj=None, rel=None, many_obj=None):
        subj = self.encode_subj(subj)
        rel = self.encode_rel(rel)
        if many_obj is None:
            many_obj = self.get_all_obj()
        return self._score(subj, rel, many_obj, prefix=True, sp=True, po=False)

    def po_prefix_score(self, rel=None, obj=None, many_subj=None):
        if many_subj is None:
            many_subj = self.get_all_subj()
        rel = self.encode_rel(rel)
        obj = self.encode_obj(obj)
        return self._score(many_subj, rel, obj, prefix=True, sp=False, po=True)

    def precompute_batch_shared_inputs(self, entity_ids):
        return self.encode_obj(entity_ids)

    # folded query vectors: score[b, n] = <q[b], E[n]>
    def sp_prefix_query(self, subj: torch.Tensor, rel: torch.Tensor) -> torch.Tensor:
        return Fn.FoldQuery.apply(self.fold_sp, _flat2d(subj), _flat2d(rel))

    def po_prefix_query(self, rel: torch.Tensor, obj: torch.Tensor) -> torch.Tensor:
        return Fn.FoldQuery.apply(self.fold_po, _flat2d(obj), _flat2d(rel))

    def _score(self, subj, rel, obj, prefix=False, sp=None, po=None, **kwargs):
        batch_sz = rel.size(0)
        subj, rel, obj = _flat2d(subj), _flat2d(rel), _flat2d(obj)
        if prefix:
            if sp:
                out = Fn.ScoreMatrix.apply(self.sp_prefix_query(subj, rel), obj)
            elif po:
                out = Fn.ScoreMatrix.apply(self.po_prefix_query(rel, obj), subj)
            else:
                raise Exception("prefix scoring needs sp=True or po=True")
        else:
            out = self._triple(subj, rel, obj)
        return out.reshape(batch_sz, -1)

    def _triple(self, subj, rel, obj):
        raise NotImplementedError


class ComplexRelationScorer(RelationScorer):
    """ComplEx (openkge/model.py:176-240): first half of the width real, second half imaginary."""

    fold_sp = FOLD_COMPLEX_SP
    fold_po = FOLD_COMPLEX_PO

    def _triple(self, subj, rel, obj):
        # Hadamard form of the non-prefix branch (openkge/model.py:231-238); not on the hot path
        r1, r2 = rel.chunk(2, dim=1)
        o1, o2 = obj.chunk(2, dim=1)
        subj_all = torch.cat((subj, subj), dim=1)
        rel_all = torch.cat((r1, rel, -r2), dim=1)
        obj_all = torch.cat((obj, o2, o1), dim=1)
        return (subj_all * obj_all * rel_all).sum(dim=1)


class DistmultRelationScorer(RelationScorer):
    """DistMult (openkge/model.py:243-278)."""

    def _triple(self, subj, rel, obj):
        return (subj * obj * rel).sum(dim=1)


# ---------------------------------------------------------------------------------------------
# embedders
# ---------------------------------------------------------------------------------------------

class RelationEmbedder(RelationModel):
    """openkge/model.py:80-139."""

    def encode_subj(self, subj) -> torch.Tensor:
        raise NotImplementedError

    def encode_rel(self, rel) -> torch.Tensor:
        raise NotImplementedError

    def encode_obj(self, obj) -> torch.Tensor:
        raise NotImplementedError

    def get_all_subj(self) -> torch.Tensor:
        raise NotImplementedError

    def get_all_rel(self) -> torch.Tensor:
        raise NotImplementedError

    def get_all_obj(self) -> torch.Tensor:
        raise NotImplementedError

    def precompute_embeddings_from_tokens(self):
        raise NotImplementedError

    # ---- batch API of the fused loss module ----
    grad_pad_rows = 0

    def encode_prefix_batch(self, po_input, sp_input, candidate_ids: Optional[torch.Tensor]):
        """Encodes everything one batch needs, in the reference's call order (openkge/trainer.py:69-87):
        candidates first (all entities when ``candidate_ids`` is None, else the batch-shared ids), then
        po: rel, obj; then sp: subj, rel. Returns (E, (rel_po, obj_po) | None, (subj_sp, rel_sp) | None)."""
        if candidate_ids is None:
            E = self.get_all_obj() if not self.training else self.encode_all_entities()
        else:
            E = _flat2d(self.precompute_batch_shared_inputs(candidate_ids.reshape(-1)))
        po = sp = None
        if po_input is not None:
            po = (self.encode_rel(po_input[0]), self.encode_obj(po_input[1]))
        if sp_input is not None:
            sp = (self.encode_subj(sp_input[0]), self.encode_rel(sp_input[1]))
        return E, po, sp

    def encode_all_entities(self) -> torch.Tensor:
        """Training-mode encode of every real entity: precompute_batch_shared_inputs(arange(2, size))
        of the 1-vs-all branch (openkge/trainer.py:80-82, openkge/dataset.py:872)."""
        raise NotImplementedError

    def encode_queries(self, po_input, sp_input, candidate_ids: Optional[torch.Tensor]):
        """(E, Q): the candidate matrix and the folded query rows of the batch, po rows first (openkge/trainer.py:69-91)."""
        E, po, sp = self.encode_prefix_batch(po_input, sp_input, candidate_ids)
        qs = []
        if po is not None:
            qs.append(self.po_prefix_query(po[0], po[1]))
        if sp is not None:
            qs.append(self.sp_prefix_query(sp[0], sp[1]))
        return E, (qs[0] if len(qs) == 1 else torch.cat(qs))


class LookupBaseRelationEmbedder(RelationEmbedder):
    """openkge/model.py:353-542. Same keyword arguments and defaults."""

    def __init__(self, entity_slot_size, relation_slot_size, train_data: EntityRelationDatasetMeta,
                 entity_embedding_size=None, relation_embedding_size=None, normalize='', dropout=0.0,
                 input_dropout=0.0, relation_dropout=0.0, relation_input_dropout=0.0, project_entity=False,
                 project_entity_activation='ReLU', project_relation=True, project_relation_activation=None,
                 sparse=False, init_std=0.01, batch_norm=False, l2_reg=0):
        super().__init__()
        self.train_data = train_data
        if relation_slot_size is None or relation_slot_size <= 0:
            relation_slot_size = entity_slot_size
        self._entity_embedding_size = entity_embedding_size if entity_embedding_size is not None else entity_slot_size
        self._relation_embedding_size = (relation_embedding_size if relation_embedding_size is not None
                                         else relation_slot_size)
        # nn.Embedding only as the parameter container (state-dict keys entity_embedding.weight, ...); lookups go through
        # the native gather. `sparse=True` (openkge/model.py:390-391): a table whose only use in a step are row lookups
        # (the relation table; the entity table with batch-shared candidates) hands the optimizer a sparse gradient
        # (functional.RowsGrad -> okge_adagrad_rows, touched rows only; like torch, Adagrad takes it only without weight
        # decay and Adam not at all). The 1-vs-all candidate gradient touches every row and stays dense / factored.
        self.entity_embedding = torch.nn.Embedding(train_data.entities_size, self._entity_embedding_size, padding_idx=PAD)
        self.relation_embedding = torch.nn.Embedding(train_data.relations_size, self._relation_embedding_size,
                                                     padding_idx=PAD)
        self.sparse = bool(sparse)
        if self.sparse:
            self.entity_embedding.weight._okge_sparse = True
            self.relation_embedding.weight._okge_sparse = True
        if project_relation:
            act = getattr(torch.nn, project_relation_activation)() if project_relation_activation else None
            lin = torch.nn.Linear(self._relation_embedding_size, entity_slot_size ** 2, bias=False)
            torch.nn.init.xavier_normal_(lin.weight.data)
            self.relation_projection = _Sequential(lin, act)
        if project_entity:
            act = getattr(torch.nn, project_entity_activation) if project_entity_activation else None
            ls = torch.nn.Linear(entity_slot_size, entity_slot_size, bias=False)
            lo = torch.nn.Linear(entity_slot_size, entity_slot_size, bias=False)
            torch.nn.init.xavier_normal_(ls.weight.data)
            torch.nn.init.xavier_normal_(lo.weight.data)
            self.subj_projection = _Sequential(ls, act() if act else None)
            self.obj_projection = _Sequential(lo, act() if act else None)
        self.project_entity = project_entity
        self.project_relation = project_relation
        self.slot_size = entity_slot_size
        self.normalize = normalize
        torch.nn.init.normal_(self.entity_embedding.weight.data, std=init_std)
        torch.nn.init.normal_(self.relation_embedding.weight.data, std=init_std)
        self.dropout = dropout
        self.input_dropout = input_dropout
        self.relation_dropout = dropout if relation_dropout is None else relation_dropout
        self.relation_input_dropout = input_dropout if relation_input_dropout is None else relation_input_dropout
        self.batch_norm = batch_norm
        if self.batch_norm:
            self.bn_e = torch.nn.BatchNorm1d(self._entity_embedding_size)
            self.bn_r = torch.nn.BatchNorm1d(self._relation_embedding_size)
        self.l2_reg = l2_reg
        self._l2_reg_hook = None
        self.grad_pad_rows = train_data.min_entities_size

    # ---- entity table partitioned over the ranks of a torch.distributed job ----------------------------------------------
    _shard = None

    def shard_entities(self, rank: int, world: int, comm=None) -> None:
        """Keeps only this rank's contiguous block of the real entity rows (``sharded.shard_bounds``; the special rows
        PAD / UNK stay in front, owned by rank 0) in ``entity_embedding.weight``: every rank scores the batch's queries
        against its block (1-vs-all), the loss sums, dQ and the rank counts are all-reduced (``functional.AllReduceSum`` /
        ``ReplicatedInput``, ``sharded.sharded_rank_counts``), the optimizer step of the table is local. Everything else
        (relation table, dropout draws of the query rows, optimizer regime) is replicated computation: ranks must be
        seeded alike and fed the same batches. Same class, same ``Trainer``; state dicts then hold the local block
        (``gather_entity_table`` reassembles the reference's tensor)."""
        from .sharded import EntityShard, _Comm, shard_bounds
        if self.batch_norm:
            raise NotImplementedError("batch norm over a sharded candidate table needs cross-rank statistics")
        if self._shard is not None:
            raise RuntimeError("the entity table is sharded already")
        ms = self.train_data.min_entities_size
        lo, hi = shard_bounds(self.train_data.entities_size - ms, world, rank)
        w = self.entity_embedding.weight
        w.data = torch.cat([w.data[:ms], w.data[ms + lo:ms + hi]]).clone()
        w.grad = None
        w._okge_shadow = None
        self._shard = EntityShard(lo, hi, int(rank), int(world), comm if comm is not None else _Comm())

    def gather_entity_table(self) -> torch.Tensor:
        """The full [entities_size, D] table on every rank (checkpoints in the reference's format)."""
        w = self.entity_embedding.weight.data
        sh = self._shard
        if sh is None:
            return w.clone()
        ms = self.train_data.min_entities_size
        full = torch.zeros((self.train_data.entities_size, w.size(1)), dtype=w.dtype, device=w.device)
        full[ms + sh.lo:ms + sh.hi] = w[ms:]
        if sh.rank == 0:
            full[:ms] = w[:ms]
        return sh.comm.all_reduce(full)

    def _entity_lookup(self, ids: torch.Tensor, with_candidates: bool = False):
        """rows = entity_embedding.weight[ids] for GLOBAL ids (and, ``with_candidates``, the training-mode 1-vs-all operand
        from the same autograd node). Sharded table: the local rows, zeros for rows of other ranks, summed over ranks."""
        w, ms, sh = self.entity_embedding.weight, self.train_data.min_entities_size, self._shard
        own = None
        if sh is not None:
            g = ids.reshape(-1).long()
            own = (g >= ms + sh.lo) & (g < ms + sh.hi)
            local = torch.where(own, g - sh.lo, torch.zeros_like(g))
            if sh.rank == 0:
                special = g < ms
                own, local = own | special, torch.where(special, g, local)
            ids = local.to(torch.int32)                          # rows of other ranks read PAD and are masked below
        e_raw = None
        if with_candidates:
            e_raw, rows = Fn.LookupAll.apply(w, ids, ms)
        else:
            rows = Fn.GatherRows.apply(w, ids.reshape(-1), PAD)
        if sh is not None:
            rows = Fn.AllReduceSum.apply(rows * own.unsqueeze(1).to(rows.dtype), sh.comm)
        return e_raw, rows

    def after_batch_loss_hook(self, epoch):
        if self.training and self.l2_reg > 0:
            result, self._l2_reg_hook = self._l2_reg_hook, None
            return result
        return None

    # _encode of the reference (openkge/model.py:455-480), split into lookup and post-processing
    def _post(self, repr, project, input_dropout, dropout, batch_norm, seg=None, n_seg=1, segment_rows=None):
        repr = self._dropout(repr, input_dropout)
        if self.batch_norm:
            repr = Fn.batch_norm_rows(batch_norm, _flat2d(repr), seg, n_seg, segment_rows)
        if project:
            repr = project(repr)
        if self.normalize == 'norm':
            repr = F.normalize(repr)
        repr = self._dropout(repr, dropout)
        if self.training and self.l2_reg > 0:
            h = repr / self.dropout if self.dropout > 0 else repr
            h = self.l2_reg * h.abs().pow(3).sum()
            self._l2_reg_hook = h if self._l2_reg_hook is None else self._l2_reg_hook + h
        return repr

    def _encode(self, slot_item, embedding, project, input_dropout, dropout, batch_norm=None, lookup=True):
        if lookup and embedding is self.entity_embedding and self._shard is not None:
            repr = self._entity_lookup(slot_item)[1]
        elif lookup:
            repr = Fn.GatherRows.apply(embedding.weight, slot_item.reshape(-1), PAD)
        else:
            repr = slot_item
        return self._post(repr, project, input_dropout, dropout, batch_norm)

    def _rel_args(self):
        return (self.relation_projection if self.project_relation else None, self.relation_input_dropout,
                self.relation_dropout, self.bn_r if self.batch_norm else None)

    def _subj_args(self):
        return (self.subj_projection if self.project_entity else None, self.input_dropout, self.dropout,
                self.bn_e if self.batch_norm else None)

    def _obj_args(self):
        return (self.obj_projection if self.project_entity else None, self.input_dropout, self.dropout,
                self.bn_e if self.batch_norm else None)

    def encode_rel(self, rel, lookup=True):
        return self._encode(rel, self.relation_embedding, *self._rel_args(), lookup=lookup)

    def encode_subj(self, subj, lookup=True):
        return self._encode(subj, self.entity_embedding, *self._subj_args(), lookup=lookup)

    def encode_obj(self, obj, lookup=True):
        return self._encode(obj, self.entity_embedding, *self._obj_args(), lookup=lookup)

    def _get_all(self, min_size, encode_func, embedding):
        # weight[min_size:] is contiguous already; no copy (openkge/model.py:512-514)
        return encode_func(embedding.weight[min_size:], lookup=False)

    def get_all_rel(self):
        return self._get_all(self.train_data.min_relations_size, self.encode_rel, self.relation_embedding)

    def get_all_subj(self):
        return self._get_all(self.train_data.min_entities_size, self.encode_subj, self.entity_embedding)

    def get_all_obj(self):
        return self._get_all(self.train_data.min_entities_size, self.encode_obj, self.entity_embedding)

    def _get(self, encode_func, id):
        dev = self.entity_embedding.weight.device
        return encode_func(torch.tensor([id], dtype=torch.int32, device=dev))

    def get_subj(self, subj):
        return self._get(self.encode_subj, subj)

    def get_rel(self, rel):
        return self._get(self.encode_rel, rel)

    def get_obj(self, obj):
        return self._get(self.encode_obj, obj)

    def get_slot_size(self):
        return self.slot_size

    def scoring_operand(self, E, split=False):
        """1-vs-all over the raw table (no dropout / batch norm / projection between the parameter and the scoring pass):
        the table's fp16 shadow copy (``functional.TableShadow``), which the fused optimizer step keeps current."""
        w = self.entity_embedding.weight
        ms = self.train_data.min_entities_size
        if (E.dim() == 2 and E.size(0) == w.size(0) - ms and E.stride(0) == w.size(1)
                and E.data_ptr() == w.data_ptr() + ms * w.size(1) * 4):
            return Fn.table_operand(w, ms, split)
        return None

    def encode_all_entities(self):
        e_all, _ = Fn.LookupAll.apply(self.entity_embedding.weight, torch.zeros(0, dtype=torch.int32,
                                      device=self.entity_embedding.weight.device), self.train_data.min_entities_size)
        return self._post(e_all, *self._obj_args())

    def encode_queries(self, po_input, sp_input, candidate_ids):
        """Without batch norm / projections the four per-block encodes of the reference (po rel, po obj, sp subj, sp rel,
        openkge/trainer.py:69-87) differ only in their dropout draws, so the entity rows and the relation rows of the
        whole batch are post-processed in one call each and folded from one autograd node (``FoldQuerySplit``): no slice,
        cat or gradient-accumulation kernels between the lookups and the scoring pass."""
        if self.project_entity or self.project_relation or self.normalize == 'norm':
            return super().encode_queries(po_input, sp_input, candidate_ids)
        e_raw, rows, rel_rows, b_po = self._lookup_batch(po_input, sp_input, candidate_ids)
        E = self._encode_candidates(e_raw, candidate_ids)
        self._candidates_are_raw_table = bool(candidate_ids is None and self.training and E is e_raw)
        seg_args = ()
        if self.batch_norm:        # statistics per block (po, sp), in the reference's call order, from ONE launch sequence
            B = rows.size(0)
            seg_args = (self._row_segments((0, b_po, b_po, B), rows.device), 2, (b_po, B - b_po))
        ent = self._post(rows, None, self.input_dropout, self.dropout, self.bn_e if self.batch_norm else None, *seg_args)
        rel = self._post(rel_rows, None, self.relation_input_dropout, self.relation_dropout,
                         self.bn_r if self.batch_norm else None, *seg_args)
        if self._graph_row_kinds is not None and self.fold_po != self.fold_sp:
            return E, Fn.FoldQueryRows.apply(self._graph_row_kinds, ent, rel)    # kinds are data: shape-static step
        return E, Fn.FoldQuerySplit.apply(self.fold_po, self.fold_sp, b_po, ent, rel)

    fuse_candidate_dropout = True

    def _encode_candidates(self, e_raw, candidate_ids):
        """``_post`` of the 1-vs-all candidate rows. When input dropout is all that stands between the table and the
        scoring pass, the rows are NOT dropped here: the dropout is named (``_candidate_dropout``, the same position in the
        Philox stream ``_post`` would have used) and ``AddLossModule.forward`` applies it to the table's fp16 operand
        (``okge_f16_mask_dropout``) and to the gradient (``okge_gemm_adagrad_dropout`` / ``okge_dropout``): the
        candidates stay the raw table, so the fused dE + Adagrad step applies (``_candidates_are_raw_table``) and the
        [N, D] fp32 copy of the dropped table is never written."""
        self._candidate_dropout = None
        if (self.fuse_candidate_dropout and candidate_ids is None and self.training and self.input_dropout > 0
                and not self.batch_norm and not self.project_entity and self.normalize != 'norm' and self.dropout <= 0
                and self.l2_reg <= 0 and e_raw.size(1) % 4 == 0):
            self._candidate_dropout = self._dropout_spec(self.input_dropout)
            return e_raw
        return self._post(e_raw, *self._obj_args())

    def _lookup_batch(self, po_input, sp_input, candidate_ids):
        """(candidate rows, entity rows of the batch [B, D] po first, relation rows [B, D], b_po)."""
        w = self.entity_embedding.weight
        ids, b_po = [], 0
        if po_input is not None:
            ids.append(po_input[1].reshape(-1))
            b_po = ids[0].numel()
        if sp_input is not None:
            ids.append(sp_input[0].reshape(-1))
        ent_ids = (ids[0] if len(ids) == 1 else torch.cat(ids)).to(torch.int32) if ids else \
            torch.zeros(0, dtype=torch.int32, device=w.device)
        if candidate_ids is None:
            if self.training:
                e_raw, rows = self._entity_lookup(ent_ids, with_candidates=True)
            else:
                e_raw = w[self.train_data.min_entities_size:]
                rows = self._entity_lookup(ent_ids)[1]
        else:
            if self._shard is not None:
                raise NotImplementedError("batch-shared candidate lists replicate the scoring; they are not sharded")
            e_raw = Fn.GatherRows.apply(w, candidate_ids.reshape(-1), PAD)
            rows = Fn.GatherRows.apply(w, ent_ids, PAD)
        rel_ids = [x[i].reshape(-1) for x, i in ((po_input, 0), (sp_input, 1)) if x is not None]
        rel_rows = Fn.GatherRows.apply(self.relation_embedding.weight, rel_ids[0] if len(rel_ids) == 1 else torch.cat(rel_ids), PAD)
        return e_raw, rows, rel_rows, b_po

    def encode_prefix_batch(self, po_input, sp_input, candidate_ids):
        """One LookupAll node for the entity table (candidates + the batch's obj/subj rows), one GatherRows
        for the relation table; post-processing in the reference's call order."""
        w = self.entity_embedding.weight
        ids, b_po = [], 0
        if po_input is not None:
            ids.append(po_input[1].reshape(-1))
            b_po = ids[0].numel()
        if sp_input is not None:
            ids.append(sp_input[0].reshape(-1))
        ent_ids = torch.cat(ids).to(torch.int32) if ids else torch.zeros(0, dtype=torch.int32, device=w.device)
        if candidate_ids is None:
            if self.training:
                e_raw, rows = self._entity_lookup(ent_ids, with_candidates=True)
            else:
                e_raw = w[self.train_data.min_entities_size:]
                rows = self._entity_lookup(ent_ids)[1]
        else:
            if self._shard is not None:
                raise NotImplementedError("batch-shared candidate lists replicate the scoring; they are not sharded")
            e_raw = Fn.GatherRows.apply(w, candidate_ids.reshape(-1), PAD)
            rows = Fn.GatherRows.apply(w, ent_ids, PAD)
        rel_ids = [x[i].reshape(-1) for x, i in ((po_input, 0), (sp_input, 1)) if x is not None]
        rel_rows = Fn.GatherRows.apply(self.relation_embedding.weight, torch.cat(rel_ids), PAD)
        E = self._encode_candidates(e_raw, candidate_ids)
        # the candidate operand IS the parameter table (no dropout / BN / projection in between): its gradient may be
        # left in factored form for the fused dE + Adagrad step (functional.DeferredTableGrad)
        self._candidates_are_raw_table = bool(candidate_ids is None and self.training and E is e_raw)
        po = sp = None
        if po_input is not None:
            po = (self._post(rel_rows[:b_po], *self._rel_args()), self._post(rows[:b_po], *self._obj_args()))
        if sp_input is not None:
            sp = (self._post(rows[b_po:], *self._subj_args()), self._post(rel_rows[b_po:], *self._rel_args()))
        return E, po, sp


class LookupSimpleRelationEmbedder(LookupBaseRelationEmbedder):
    """openkge/model.py:545-558."""

    def __init__(self, entity_slot_size, **kwargs):
        kwargs.pop('relation_slot_size', None)
        super().__init__(entity_slot_size=entity_slot_size, relation_slot_size=entity_slot_size,
                         project_relation=False, **kwargs)
        self.relation_projection = None


class TokenBasedRelationEmbedder(RelationEmbedder):
    """openkge/model.py:561-712: entity / relation ids -> padded token-id rows -> token embeddings."""

    def __init__(self, train_data: EntityRelationDatasetMeta, entity_slot_size: int, relation_slot_size: int,
                 sparse: bool, init_std: float, normalize=None):
        super().__init__()
        if relation_slot_size is None or relation_slot_size <= 0:
            relation_slot_size = entity_slot_size
        self.train_data = train_data
        ent_len, rel_len = train_data.max_length[0], train_data.max_length[1]
        # int64 buffers under the reference's names (state-dict compatible) + int32 copies for the kernels
        self.register_buffer('entity_token_ids', self._token_rows(train_data.entity_id_to_tokens_map, ent_len))
        self.register_buffer('relation_token_ids', self._token_rows(train_data.relation_id_to_tokens_map, rel_len))
        self.register_buffer('_entity_token_ids_i32', self.entity_token_ids.to(torch.int32), persistent=False)
        self.register_buffer('_relation_token_ids_i32', self.relation_token_ids.to(torch.int32), persistent=False)
        self.entity_embedding = torch.nn.Embedding(train_data.entity_tokens_size, entity_slot_size, padding_idx=0)
        self.relation_embedding = torch.nn.Embedding(train_data.relation_tokens_size, relation_slot_size, padding_idx=0)
        self.entity_batchnorm = None
        self.relation_batchnorm = None
        self.normalize = normalize
        if normalize == 'batchnorm':
            self.entity_batchnorm = torch.nn.BatchNorm1d(entity_slot_size, momentum=0.1, eps=1e-5)
            self.relation_batchnorm = torch.nn.BatchNorm1d(relation_slot_size, momentum=0.1, eps=1e-5)
            torch.nn.init.uniform_(self.entity_batchnorm.weight)
            torch.nn.init.uniform_(self.relation_batchnorm.weight)
        self.entity_embedding_from_tokens = None
        self.relations_embedding_from_tokens = None
        self.slot_size = entity_slot_size
        self.relation_slot_size = relation_slot_size
        # the PAD row is overwritten by the init exactly like the reference (openkge/model.py:633-634)
        torch.nn.init.normal_(self.entity_embedding.weight.data, std=init_std)
        torch.nn.init.normal_(self.relation_embedding.weight.data, std=init_std)

    @staticmethod
    def _token_rows(id_to_tokens, max_len: int) -> torch.Tensor:
        """Last ``max_len`` tokens of every row, left-aligned, PAD-filled (openkge/model.py:579-595). An int array
        that already has this layout ([rows, max_len], e.g. a decoded cache or a synthetic graph) is used as is."""
        if hasattr(id_to_tokens, "shape") and len(id_to_tokens.shape) == 2 and id_to_tokens.shape[1] == max_len:
            return torch.as_tensor(id_to_tokens).to(torch.int64).contiguous()
        rows = torch.zeros(len(id_to_tokens), max_len, dtype=torch.int64)
        for i, toks in enumerate(id_to_tokens):
            t = list(toks)[-max_len:]
            if t:
                rows[i, :len(t)] = torch.tensor(t, dtype=torch.int64)
        return rows

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        # in place: captured CUDA graphs (graphed.GraphedTrainStep) hold the addresses of these buffers
        self._entity_token_ids_i32.copy_(self.entity_token_ids)
        self._relation_token_ids_i32.copy_(self.relation_token_ids)
        self._reset_cache()
        return out

    def _all_entity_ids(self, device) -> torch.Tensor:
        """ids of every real entity (arange(min_entities_size, entities_size), openkge/dataset.py:872), int32, cached."""
        ids = self.__dict__.get("_all_ids")
        if ids is None or ids.device != device:
            ids = torch.arange(self.train_data.min_entities_size, self.entity_token_ids.size(0), dtype=torch.int32, device=device)
            self.__dict__["_all_ids"] = ids
        return ids

    def _reset_cache(self):
        self.entity_embedding_from_tokens = None
        self.relations_embedding_from_tokens = None
        self.__dict__["_eval_cache_f16"] = None

    def scoring_operand(self, E, split=False):
        """Evaluation against the cached encode of every entity: its fp16 operand is built once per cache."""
        cache = self.entity_embedding_from_tokens
        ms = self.train_data.min_entities_size
        if (cache is None or self.training or E.dim() != 2 or E.size(0) != cache.size(0) - ms
                or E.data_ptr() != cache.data_ptr() + ms * cache.size(1) * 4):
            return None
        op = self.__dict__.get("_eval_cache_f16")
        if op is None or (split and op.lo is None):
            from . import kernels as K
            op = K.quantize(cache[ms:], split=split)
            self.__dict__["_eval_cache_f16"] = op
        return op if split else op.without_lo()

    def eval(self, *args, **kwargs):
        self._reset_cache()
        return super().eval()

    def train(self, *args, **kwargs):
        self._reset_cache()
        return super().train(*args, **kwargs)

    def get_all_subj(self):
        self.precompute_embeddings_from_tokens()
        return self.entity_embedding_from_tokens[self.train_data.min_entities_size:]

    def get_all_obj(self):
        self.precompute_embeddings_from_tokens()
        return self.entity_embedding_from_tokens[self.train_data.min_entities_size:]

    def get_all_rel(self):
        self.precompute_embeddings_from_tokens()
        return self.relations_embedding_from_tokens[self.train_data.min_relations_size:]

    def get_subj(self, subj):
        self.precompute_embeddings_from_tokens()
        return self.entity_embedding_from_tokens[subj].unsqueeze(0)

    def get_rel(self, rel):
        self.precompute_embeddings_from_tokens()
        return self.relations_embedding_from_tokens[rel]

    def get_obj(self, obj):
        self.precompute_embeddings_from_tokens()
        return self.entity_embedding_from_tokens[obj].unsqueeze(0)

    def precompute_embeddings_from_tokens(self):
        """Eval-mode encode of EVERY entity and relation row, cached until the next train()/eval()
        (openkge/model.py:670-712). One pooled-gather launch per table on the device instead of the
        reference's 4,096-row chunks through a CPU tensor; like the reference it leaves the module in
        eval mode (:682)."""
        if self.entity_embedding_from_tokens is not None:
            return
        super().eval()
        with torch.no_grad():
            self.entity_embedding_from_tokens = _flat2d(self._encode_rows('entity', None, 0, self.entity_token_ids.size(0)))
            self.relations_embedding_from_tokens = _flat2d(
                self._encode_rows('relation', None, 0, self.relation_token_ids.size(0)))

    def _encode_rows(self, which, ids, id_start=0, n=None, seg=None, n_seg=1, segment_rows=None, zero_tail=False):
        raise NotImplementedError

    def encode_subj(self, subj):
        return self._encode_rows('entity', subj.reshape(-1))

    def encode_obj(self, obj):
        return self._encode_rows('entity', obj.reshape(-1))

    def encode_rel(self, rel):
        return self._encode_rows('relation', rel.reshape(-1))

    def encode_all_entities(self):
        lo = self.train_data.min_entities_size
        return _flat2d(self._encode_rows('entity', None, lo, self.entity_token_ids.size(0) - lo))

    def get_slot_size(self):
        return self.slot_size

    def encode_queries(self, po_input, sp_input, candidate_ids):
        """The four per-block encodes of the reference (po rel, po obj, sp subj, sp rel, openkge/trainer.py:69-87) as one
        encode call per table over the rows of the whole batch; batch-norm statistics stay per block (row segments
        (0, b_po, B) handed to the kernels as device data) and update the running statistics in the reference's order
        (candidates, po block, sp block). The folded query rows come from one autograd node."""
        if getattr(self, 'relation_projection', None) is not None or getattr(self, 'entity_projection', None) is not None:
            return super().encode_queries(po_input, sp_input, candidate_ids)
        ent_ids = [x[i].reshape(-1) for x, i in ((po_input, 1), (sp_input, 0)) if x is not None]
        rel_ids = [x[i].reshape(-1) for x, i in ((po_input, 0), (sp_input, 1)) if x is not None]
        b_po = 0 if po_input is None else po_input[0].numel()
        ent_ids = ent_ids[0] if len(ent_ids) == 1 else torch.cat(ent_ids)
        rel_ids = rel_ids[0] if len(rel_ids) == 1 else torch.cat(rel_ids)
        B = ent_ids.numel()
        dev = ent_ids.device
        train_bn = self.normalize == 'batchnorm' and self.training
        graphed = self._graph_segments is not None
        if candidate_ids is None and not self.training:
            E = self.get_all_obj()                                       # cached eval-mode encode of every entity
            ent = _flat2d(self._encode_rows('entity', ent_ids))
        else:
            # candidates and the batch's own entity rows go through the encoder in ONE call (one gradient buffer for the
            # token table, no dense accumulation pass); the split is undone by row ranges
            if candidate_ids is None:
                cand = self._all_entity_ids(dev)
            else:
                cand = candidate_ids.reshape(-1)
            N = cand.numel()
            padded = self._graph_candidate_count is not None            # rows [count, N) are padding (graph replay)
            seg_args = {}
            if train_bn:
                seg_args = dict(seg=self._row_segments((0, N, N, N + b_po, N + b_po, N + B), dev), n_seg=3,
                                segment_rows=None if graphed else (N, b_po, B - b_po), zero_tail=padded)
            both = _flat2d(self._encode_rows('entity', torch.cat([cand.to(ent_ids.dtype), ent_ids]), **seg_args))
            E, ent = Fn.SplitRows.apply(both, N)
        seg_args = {}
        if train_bn:
            seg = self._graph_segments[6:10] if graphed else self._row_segments((0, b_po, b_po, B), dev)
            seg_args = dict(seg=seg, n_seg=2, segment_rows=None if graphed else (b_po, B - b_po))
        rel = _flat2d(self._encode_rows('relation', rel_ids, **seg_args))
        if self._graph_row_kinds is not None and self.fold_po != self.fold_sp:
            return E, Fn.FoldQueryRows.apply(self._graph_row_kinds, ent, rel)
        return E, Fn.FoldQuerySplit.apply(self.fold_po, self.fold_sp, b_po, ent, rel)



class UnigramPoolingRelationEmbedder(TokenBasedRelationEmbedder):
    """openkge/model.py:716-798. ``entity_projection`` is defined here (``None``): the reference reads
    it without ever assigning it (:789, :792), so its own forward raises AttributeError."""

    def __init__(self, entity_slot_size, relation_slot_size, train_data: EntityRelationDatasetMeta, pool='sum',
                 normalize=None, dropout=0.0, entity_dropout=None, relation_dropout=None, sparse=False,
                 init_std=0.01, activation=None, project_relation=False):
        super().__init__(entity_slot_size=entity_slot_size, relation_slot_size=relation_slot_size,
                         train_data=train_data, sparse=sparse, init_std=init_std, normalize=normalize)
        if relation_slot_size is None or relation_slot_size <= 0:
            relation_slot_size = entity_slot_size
        self.relation_slot_size = relation_slot_size
        self.relation_projection = None
        self.entity_projection = None
        if project_relation:
            self.relation_slot_size = entity_slot_size ** 2
            lin = torch.nn.Linear(relation_slot_size, entity_slot_size ** 2, bias=False)
            torch.nn.init.normal_(lin.weight.data, 1 / (entity_slot_size ** 2 * relation_slot_size * init_std ** 3))
            self.relation_projection = _Sequential(lin, torch.nn.BatchNorm1d(entity_slot_size ** 2))
        self.pool = pool if pool in ('max', 'mean') else 'sum'
        self.entity_dropout = entity_dropout if entity_dropout else dropout
        self.relation_dropout = relation_dropout if relation_dropout else dropout
        self.activation = getattr(torch.nn, activation)() if activation is not None and hasattr(torch.nn, activation) else None
        self.grad_pad_rows = 0

    def _encode_rows(self, which, ids, id_start=0, n=None, seg=None, n_seg=1, segment_rows=None, zero_tail=False):
        if which == 'entity':
            emb, rows, proj, p, norm = (self.entity_embedding, self._entity_token_ids_i32, self.entity_projection,
                                        self.entity_dropout, self.entity_batchnorm)
        else:
            emb, rows, proj, p, norm = (self.relation_embedding, self._relation_token_ids_i32,
                                        self.relation_projection, self.relation_dropout, self.relation_batchnorm)
        encoded = Fn.GatherPool.apply(emb.weight, rows, ids, self.pool, id_start, n)     # :763-774
        if self.activation is not None:
            encoded = self.activation(encoded)
        if self.normalize == 'norm':
            encoded = F.normalize(encoded, dim=1)
        if self.normalize == 'batchnorm':
            if self.training and p > 0 and not proj:
                # batch norm directly followed by dropout: one fused pass (the mask is drawn in the kernel's epilogue)
                return Fn.batch_norm_rows(norm, encoded, seg, n_seg, segment_rows, zero_tail,
                                          dropout=self._dropout_spec(p)).unsqueeze(1)
            encoded = Fn.batch_norm_rows(norm, encoded, seg, n_seg, segment_rows, zero_tail)   # :777-780
        if proj:
            encoded = proj(encoded)
        return self._dropout(encoded, p).unsqueeze(1)                                    # :783-786


class LSTMRelationEmbedder(TokenBasedRelationEmbedder):
    """openkge/model.py:912-998: tokens -> embeddings -> single-layer LSTM -> hidden state at the last real token ->
    [batch norm] -> [relation projection] -> dropout. ``entity_encoder_in`` / ``relation_encoder_in`` are
    ``torch.nn.LSTM`` modules used as PARAMETER CONTAINERS only (state-dict keys ``*_encoder_in.weight_ih_l0`` ...): the
    recurrence itself runs on the native kernels (``functional.LSTMLastState``), not on cuDNN.

    Differences to the reference, both in its favour: ``encoder_activiation`` is applied as an activation (the reference
    instantiates the class with the tensor as constructor argument, :976, which cannot work), and rows are encoded in
    chunks of ``chunk_rows`` so that the all-entities passes fit in memory."""

    chunk_rows = 1 << 15

    def __init__(self, entity_slot_size, relation_slot_size, train_data: EntityRelationDatasetMeta, dropout=0.0,
                 entity_dropout=None, relation_dropout=None, encoder_activiation=None, sparse=False, init_std=0.1,
                 normalize='', project_relation=False):
        super().__init__(entity_slot_size=entity_slot_size, relation_slot_size=relation_slot_size,
                         train_data=train_data, sparse=sparse, init_std=init_std, normalize=normalize)
        if relation_slot_size is None or relation_slot_size <= 0:
            relation_slot_size = entity_slot_size
        self.relation_slot_size = relation_slot_size
        self.relation_projection = None
        self.entity_projection = None
        if project_relation:
            self.relation_slot_size = entity_slot_size ** 2
            lin = torch.nn.Linear(relation_slot_size, entity_slot_size ** 2, bias=False)
            torch.nn.init.normal_(lin.weight.data, 1 / (entity_slot_size ** 2 * relation_slot_size * init_std ** 3))
            self.relation_projection = _Sequential(lin, torch.nn.BatchNorm1d(entity_slot_size ** 2))
        self.encoder_activiation = None
        if encoder_activiation is not None and hasattr(torch.nn, encoder_activiation):
            self.encoder_activiation = getattr(torch.nn, encoder_activiation)()
        self.entity_encoder_in = torch.nn.LSTM(input_size=entity_slot_size, hidden_size=entity_slot_size, batch_first=True)
        self.relation_encoder_in = torch.nn.LSTM(input_size=relation_slot_size, hidden_size=relation_slot_size,
                                                 batch_first=True)
        self.entity_dropout = entity_dropout if entity_dropout else dropout
        self.relation_dropout = relation_dropout if relation_dropout else dropout
        self.grad_pad_rows = 0

    def _encode_rows(self, which, ids, id_start=0, n=None, seg=None, n_seg=1, segment_rows=None, zero_tail=False):
        if which == 'entity':
            emb, rows, enc, proj, p, norm = (self.entity_embedding, self._entity_token_ids_i32, self.entity_encoder_in,
                                             None, self.entity_dropout, self.entity_batchnorm)
        else:
            emb, rows, enc, proj, p, norm = (self.relation_embedding, self._relation_token_ids_i32,
                                             self.relation_encoder_in, self.relation_projection, self.relation_dropout,
                                             self.relation_batchnorm)
        L = rows.size(1)
        if ids is None:
            ids = torch.arange(id_start, id_start + (n if n is not None else rows.size(0) - id_start), device=rows.device)
        ids = ids.reshape(-1).long()
        parts = []
        for lo in range(0, ids.numel(), self.chunk_rows):
            tok = rows.index_select(0, ids[lo:lo + self.chunk_rows])                     # _map_to_tokens, :957-961
            last = ((tok > 0).sum(1) - 1).remainder(L).to(torch.int32)                    # :970 (-1 wraps like the reference)
            parts.append(Fn.LSTMLastState.apply(emb.weight, enc.weight_ih_l0, enc.weight_hh_l0, enc.bias_ih_l0,
                                                enc.bias_hh_l0, tok.t().contiguous(), last))
        encoded = parts[0] if len(parts) == 1 else torch.cat(parts)
        if self.encoder_activiation is not None:
            encoded = self.encoder_activiation(encoded)
        if self.normalize == 'batchnorm':
            if self.training and p > 0 and proj is None:
                return Fn.batch_norm_rows(norm, encoded, seg, n_seg, segment_rows, zero_tail,
                                          dropout=self._dropout_spec(p)).unsqueeze(1)       # fused :981-986
            encoded = Fn.batch_norm_rows(norm, encoded, seg, n_seg, segment_rows, zero_tail)   # :981-982
        if proj is not None:
            encoded = proj(encoded)
        if p > 0:
            return self._dropout(encoded, p).unsqueeze(1)                                # :985-986
        return encoded                                                                   # :987-988 ([n, D], no unsqueeze)


# ---------------------------------------------------------------------------------------------
# compositions and registry — openkge/model.py:1001-1066
# ---------------------------------------------------------------------------------------------

class LookupComplexRelationModel(ComplexRelationScorer, LookupSimpleRelationEmbedder):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class LookupDistmultRelationModel(DistmultRelationScorer, LookupSimpleRelationEmbedder):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class UnigramPoolingComplexRelationModel(ComplexRelationScorer, UnigramPoolingRelationEmbedder):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class UnigramPoolingDistmultRelationModel(DistmultRelationScorer, UnigramPoolingRelationEmbedder):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class LSTMComplexRelationModel(ComplexRelationScorer, LSTMRelationEmbedder):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class LSTMDistmultRelationModel(DistmultRelationScorer, LSTMRelationEmbedder):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class Models:
    """Registry looked up by name (``getattr(Models, args["model"])``, scripts/train.py:88). Only the model
    families on the accelerated path are registered; Rescal/Tucker3, the Bigram encoder and the
    data-bias diagnostics of the reference are out of scope."""

    LookupDistmultRelationModel = LookupDistmultRelationModel
    LookupComplexRelationModel = LookupComplexRelationModel
    UnigramPoolingComplexRelationModel = UnigramPoolingComplexRelationModel
    UnigramPoolingDistmultRelationModel = UnigramPoolingDistmultRelationModel
    LSTMComplexRelationModel = LSTMComplexRelationModel
    LSTMDistmultRelationModel = LSTMDistmultRelationModel
