"""Data boundary of the hot path: the reference's batch wire format with sparse labels.

The reference collate (``OneToNMentionRelationDataset_collate_func``, openkge/dataset.py:724-940)
materialises a dense fp32 ``[B, N]`` label tensor and a dense bool ``[B, N]`` filter mask per batch —
41 GB at B = 4096, N = 2.5 M. The B200 path carries the same information as CSR index lists and keeps
the 7-tuple layout of the reference batch (openkge/dataset.py:937-940):

    (slot_inputs[2], normalizer_loss, normalizer_metric, labels, label_ids, filter_mask, shared_ids)

with ``labels`` / ``filter_mask`` as :class:`CSRMatrix` and ``label_ids`` as :class:`RankedAnswers`.
Dense tensors / list-of-lists in the reference's own format are accepted everywhere too and are
converted on the device.

Input contract = the reference's dataset tensors (openkge/dataset.py:567-710):
``seen_prefixes_tensor [P, 7] = (a, b, this_start, this_end, all_start, all_end, slot)``,
``seen_entities_tensor`` (packed ragged lists, utils/misc.py:56-89) and ``all_splits_entities_tensor``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import kernels as K
from .metrics import MetricResult

PAD, UNK, BOS, EOS = 0, 1, 2, 3  # openkge/index_mapper.py:14


@dataclass
class EntityRelationDatasetMeta:
    """Same fields as the reference dataclass (openkge/dataset.py:25-39); ``entities_size`` /
    ``relations_size`` include the two special ids, ``min_*_size`` = 2."""
    entity_id_count_map: Optional[dict] = None
    relation_id_count_map: Optional[dict] = None
    entity_token_id_count_map: Optional[dict] = None
    relation_token_id_count_map: Optional[dict] = None
    entity_id_to_tokens_map: Optional[Sequence[Sequence[int]]] = None
    relation_id_to_tokens_map: Optional[Sequence[Sequence[int]]] = None
    entities_size: int = 0
    relations_size: int = 0
    min_entities_size: int = 2
    min_relations_size: int = 2
    entity_tokens_size: int = 0
    relation_tokens_size: int = 0
    max_length: Tuple[int, int] = (10, 10)


# ---------------------------------------------------------------------------------------------
# sparse wire types
# ---------------------------------------------------------------------------------------------

class CSRMatrix:
    """0/1 matrix [B, N] as CSR (int32 row pointers + ascending unique column indices). Stands in for
    the dense ``label_tensor`` / ``filter_mask_tensor`` of the reference batch."""

    def __init__(self, ptr: torch.Tensor, idx: torch.Tensor, shape: Tuple[int, int]):
        self.ptr, self.idx, self.shape = ptr, idx, (int(shape[0]), int(shape[1]))

    @staticmethod
    def from_lists(rows: Sequence[Sequence[int]], n_cols: int) -> "CSRMatrix":
        ptr = np.zeros(len(rows) + 1, np.int32)
        ptr[1:] = np.cumsum([len(r) for r in rows])
        idx = np.concatenate([np.asarray(sorted(r), np.int32) for r in rows]) if len(rows) else np.zeros(0, np.int32)
        return CSRMatrix(torch.from_numpy(ptr), torch.from_numpy(idx.astype(np.int32)), (len(rows), n_cols))

    @staticmethod
    def from_dense(dense: torch.Tensor) -> "CSRMatrix":
        """Dense [B, N] (fp32 labels or bool mask) -> CSR, on the tensor's device. Row-major nonzero order
        gives ascending columns inside each row."""
        nz = dense != 0
        counts = nz.sum(dim=1)
        ptr = torch.zeros(dense.size(0) + 1, dtype=torch.int32, device=dense.device)
        ptr[1:] = counts.cumsum(0)
        idx = nz.nonzero(as_tuple=False)[:, 1].to(torch.int32)
        return CSRMatrix(ptr, idx, tuple(dense.shape))

    def to(self, device, non_blocking: bool = False) -> "CSRMatrix":
        return CSRMatrix(self.ptr.to(device, non_blocking=non_blocking), self.idx.to(device, non_blocking=non_blocking),
                         self.shape)

    def pin_memory(self) -> "CSRMatrix":
        return CSRMatrix(self.ptr.pin_memory(), self.idx.pin_memory(), self.shape)

    def to_dense(self, dtype=torch.float32) -> torch.Tensor:
        out = torch.zeros(self.shape, dtype=dtype, device=self.ptr.device)
        rows = torch.repeat_interleave(torch.arange(self.shape[0], device=self.ptr.device),
                                       (self.ptr[1:] - self.ptr[:-1]).long())
        out[rows, self.idx.long()] = 1
        return out

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def nnz(self) -> int:
        return int(self.idx.numel())

    def sum(self) -> float:
        return float(self.nnz)

    @property
    def nbytes(self) -> int:
        return self.ptr.numel() * 4 + self.idx.numel() * 4


class RankedAnswers:
    """The ragged ``label_ids`` of the reference eval batch (list[B] of list of IntTensor,
    openkge/dataset.py:923-926) flattened: ranked answer j belongs to prefix row ``ans_row[j]`` and has
    the alternative mention columns ``alt_idx[alt_ptr[j]:alt_ptr[j+1]]``."""

    # Ranked answers per prefix row the single-pass evaluation kernel counts (okge_score_bce_rank): 1, 2 or 4. None: chosen
    # per batch when it is collated (every slot costs four instructions per score in the kernel's epilogue, every overflow
    # answer an extra query row); a number forces it.
    SLOTS: Optional[int] = None
    LOSS_COST, SLOT_COST = 13, 4          # epilogue instructions per score: loss part, one ranking slot

    def __init__(self, ans_row: torch.Tensor, alt_ptr: torch.Tensor, alt_idx: torch.Tensor,
                 overflow: Optional[torch.Tensor] = None, overflow_slot: Optional[torch.Tensor] = None,
                 extra_prefix: Optional[torch.Tensor] = None, slots: Optional[int] = None, n_rows: Optional[int] = None):
        self.ans_row, self.alt_ptr, self.alt_idx = ans_row, alt_ptr, alt_idx
        # Computed on the HOST when the batch is collated (None = unknown, e.g. a structure built from device tensors), so
        # that the evaluation knows its launch shapes without reading device data:
        #   slots         answers a prefix row holds in the single-pass kernel
        #   overflow      int64 [n_ov]  indices of the answers beyond the first `slots` of their prefix row
        #   extra_prefix  int32 [n_x]   the single-pass kernel gets n_x extra query rows; extra row x repeats prefix row
        #                               extra_prefix[x] and holds up to `slots` of its overflow answers
        #   overflow_slot int64 [n_ov]  x * 4 + slot of every overflow answer inside the extra rows
        if overflow is None and not ans_row.is_cuda:
            r = ans_row.numpy().astype(np.int64)
            pos = np.arange(r.size) - np.searchsorted(r, r, side="left")           # ans_row is ascending (collate order)
            if slots is None:
                slots = self.SLOTS if self.SLOTS is not None else self._pick_slots(r, pos, n_rows)
            ov = np.flatnonzero(pos >= slots)
            group = (pos[ov] - slots) // slots                                     # which extra row of its prefix
            key = r[ov] * (int(pos.max()) // slots + 2 if r.size else 1) + group
            uniq, inv = np.unique(key, return_inverse=True)                        # extra rows in (prefix, group) order
            first_of = np.zeros(uniq.size, np.int64)
            first_of[inv[::-1]] = ov[::-1]
            overflow = torch.from_numpy(ov.astype(np.int64))
            overflow_slot = torch.from_numpy((inv * 4 + (pos[ov] - slots) % slots).astype(np.int64))
            extra_prefix = torch.from_numpy(r[first_of].astype(np.int32))
        self.slots = int(slots) if slots is not None else (self.SLOTS if self.SLOTS is not None else 4)
        self.overflow, self.overflow_slot, self.extra_prefix = overflow, overflow_slot, extra_prefix

    @classmethod
    def _pick_slots(cls, r: np.ndarray, pos: np.ndarray, n_rows: Optional[int]) -> int:
        """Slot count with the cheapest pass: (loss + 4 per slot) instructions per score over 128-row tiles of the prefix
        rows plus the extra rows that the answers beyond the slots need."""
        if r.size == 0:
            return 1
        rows = int(n_rows) if n_rows is not None else int(r.max()) + 1
        counts = np.bincount(r, minlength=rows)
        best, best_cost = 4, None
        for s in (1, 2, 4):
            extra = int(np.ceil(np.maximum(counts - s, 0) / s).sum())
            cost = (cls.LOSS_COST + cls.SLOT_COST * s) * -(-(rows + extra) // 128)
            if best_cost is None or cost < best_cost:
                best, best_cost = s, cost
        return best

    @staticmethod
    def from_label_ids(label_ids: Sequence[Sequence[torch.Tensor]]) -> "RankedAnswers":
        ans_row, alt_ptr, alt_idx = [], [0], []
        for b, labels in enumerate(label_ids):
            for alt in labels:
                ans_row.append(b)
                alt_idx.extend(int(x) for x in alt.reshape(-1).tolist())
                alt_ptr.append(len(alt_idx))
        return RankedAnswers(torch.tensor(ans_row, dtype=torch.int32), torch.tensor(alt_ptr, dtype=torch.int32),
                             torch.tensor(alt_idx, dtype=torch.int32))

    def to(self, device, non_blocking: bool = False) -> "RankedAnswers":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)          # noqa: E731
        return RankedAnswers(mv(self.ans_row), mv(self.alt_ptr), mv(self.alt_idx), mv(self.overflow),
                             mv(self.overflow_slot), mv(self.extra_prefix), slots=self.slots)

    def pin_memory(self) -> "RankedAnswers":
        pin = lambda t: None if t is None else t.pin_memory()                                  # noqa: E731
        return RankedAnswers(pin(self.ans_row), pin(self.alt_ptr), pin(self.alt_idx), pin(self.overflow),
                             pin(self.overflow_slot), pin(self.extra_prefix), slots=self.slots)

    def __len__(self) -> int:
        return int(self.ans_row.numel())

    @property
    def nbytes(self) -> int:
        n = 4 * (self.ans_row.numel() + self.alt_ptr.numel() + self.alt_idx.numel())
        if self.overflow is not None:
            n += 16 * self.overflow.numel() + 4 * self.extra_prefix.numel()
        return n


class AllEntityIds:
    """Symbolic ``batch_shared_entity_ids`` of 1-vs-all mode: the reference ships
    ``arange(entity_vocab_size)[offset:].int().unsqueeze(1)`` with every batch (openkge/dataset.py:872), 4 MB
    per step at 1 M entities; the B200 path only needs to know that ALL entities are the candidates."""

    def __init__(self, offset: int, n: int):
        self.offset, self.n = int(offset), int(n)
        self.shape = (self.n, 1)

    def to(self, device, non_blocking: bool = False) -> "AllEntityIds":
        return self

    def materialize(self, device="cpu") -> torch.Tensor:
        return torch.arange(self.offset, self.offset + self.n, dtype=torch.int32, device=device).unsqueeze(1)

    def view(self, *shape):
        return self.materialize().view(*shape)

    reshape = view


class PrefixScores:
    """Lazy ``all_outputs``: the [B, N] prefix scores represented by their factors (Q, E). The fused
    ranking path consumes the factors; ``dense()`` materialises the matrix the reference would return.

    ``split`` (default): the factors are contracted in split precision (fp16 hi + lo planes, three-term product): scores
    and therefore ranks are those of an fp32 scorer up to ~1e-6 norm-wise. ``e16``: the candidates' fp16 operand when
    the model keeps one (a table's shadow copy, the eval cache of a token model)."""

    def __init__(self, q: torch.Tensor, e: torch.Tensor, pending_loss: Optional[dict] = None, e16=None, split: bool = True,
                 shard=None):
        self.q, self.e = q, e
        self.shape = (q.size(0), e.size(0))
        # sharded.EntityShard when ``e`` is this rank's block of the candidates: the ranking all-reduces its counts
        self.shard = shard
        self.split = bool(split)
        self._q16 = None
        self._e16 = e16 if (e16 is None or not self.split or e16.lo is not None) else None
        # evaluation inside Trainer.compute_one_batch: the BCE loss of the batch has not been computed yet (labels, label
        # values and the [1] float64 output buffer are kept here) so that the ranking pass can produce it on the way
        self.pending_loss = pending_loss

    def operands(self):
        """(q16, e16): the fp16 operands every kernel of the ranking path contracts, quantized once per batch."""
        if self._q16 is None:
            self._q16 = K.quantize(self.q, split=self.split)
        if self._e16 is None:
            self._e16 = K.quantize(self.e, split=self.split)
        return self._q16, self._e16

    def ensure_loss(self) -> None:
        """Computes the deferred loss with its own pass if the ranking did not (no ranked answers, > 4 answers in a row)."""
        pend, self.pending_loss = self.pending_loss, None
        if pend is not None:
            q16, e16 = self.operands()
            loss, _ = K.score_bce(q16, e16, pend["ptr"], pend["idx"], pend["y_base"], pend["y_pos"], want_dS=False)
            pend["out"].copy_(loss)

    def dense(self) -> torch.Tensor:
        q16, e16 = self.operands()
        return K.score_store(q16, e16, split=self.split)

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]


# ---------------------------------------------------------------------------------------------
# dataset index + collate
# ---------------------------------------------------------------------------------------------

def _csr_take(ptr: np.ndarray, idx: np.ndarray, rows: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Row gather of a CSR structure (vectorised)."""
    lens = (ptr[rows + 1] - ptr[rows]).astype(np.int64)
    out_ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(lens, out=out_ptr[1:])
    total = int(out_ptr[-1])
    if total == 0:
        return out_ptr, np.zeros(0, idx.dtype)
    starts = np.repeat(ptr[rows].astype(np.int64) - out_ptr[:-1], lens)
    return out_ptr, idx[starts + np.arange(total, dtype=np.int64)]


def _sorted_unique_rows(row_of: np.ndarray, vals: np.ndarray, n_rows: int) -> Tuple[np.ndarray, np.ndarray]:
    """CSR of the per-row sorted unique values."""
    if len(vals) == 0:
        return np.zeros(n_rows + 1, np.int64), np.zeros(0, np.int32)
    order = np.lexsort((vals, row_of))
    r, v = row_of[order], vals[order]
    keep = np.ones(len(v), bool)
    keep[1:] = (r[1:] != r[:-1]) | (v[1:] != v[:-1])
    r, v = r[keep], v[keep]
    ptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=n_rows), out=ptr[1:])
    return ptr, v.astype(np.int32)


class PrefixIndex:
    """All prefixes of one split decoded ONCE from the reference's dataset tensors into CSR arrays, so
    that a batch is a vectorised row gather instead of the reference's per-row Python loops
    (openkge/dataset.py:762-811, 885-932). Column indices are candidate-local (entity id - offset)."""

    def __init__(self, seen_prefixes: np.ndarray, seen_entities: np.ndarray, all_splits_entities: Optional[np.ndarray],
                 entity_vocab_size: int, entity_vocab_offset: int = 2, is_training_data: bool = True):
        sp = np.asarray(seen_prefixes, dtype=np.int64).reshape(-1, 7)
        se = np.asarray(seen_entities, dtype=np.int64).reshape(-1)
        self.n_cols = int(entity_vocab_size - entity_vocab_offset)
        self.offset = int(entity_vocab_offset)
        self.is_training_data = is_training_data
        self.prefix = sp[:, :2].astype(np.int32)
        self.slot = sp[:, 6].astype(np.int32)
        P = len(sp)
        ts, te = sp[:, 2], sp[:, 3]
        # packed segment = [k+1 offsets, 0, values]; offsets[0] = header length = k + 2
        k = (se[ts] - 2) if P else np.zeros(0, np.int64)                    # answers per prefix
        self.ans_ptr = np.zeros(P + 1, np.int64)
        np.cumsum(k, out=self.ans_ptr[1:])
        Q = int(self.ans_ptr[-1])
        # alternative-list lengths: diff of the k+1 offsets of each segment
        off_pos = np.repeat(ts - self.ans_ptr[:-1], k) + np.arange(Q, dtype=np.int64)  # position of offset j
        alt_len = se[off_pos + 1] - se[off_pos] if Q else np.zeros(0, np.int64)
        self.alt_ptr = np.zeros(Q + 1, np.int64)
        np.cumsum(alt_len, out=self.alt_ptr[1:])
        n_vals = te - (ts + k + 2)                                          # values per prefix
        val_ptr = np.zeros(P + 1, np.int64)
        np.cumsum(n_vals, out=val_ptr[1:])
        V = int(val_ptr[-1])
        val_pos = np.repeat(ts + k + 2 - val_ptr[:-1], n_vals) + np.arange(V, dtype=np.int64)
        vals = (se[val_pos] - self.offset) if V else np.zeros(0, np.int64)
        self.alt_idx = vals.astype(np.int32)
        # label columns = unique answers mentions of the prefix (label_tensor[...] = 1, :921)
        self.lab_ptr, self.lab_idx = _sorted_unique_rows(np.repeat(np.arange(P), n_vals), vals, P)
        # filter columns = all-splits answers of the prefix (:927); training rows carry (0, 0)
        if all_splits_entities is not None and not is_training_data:
            ae = np.asarray(all_splits_entities, dtype=np.int64).reshape(-1)
            a0, a1 = sp[:, 4], sp[:, 5]
            n_f = a1 - a0
            f_ptr = np.zeros(P + 1, np.int64)
            np.cumsum(n_f, out=f_ptr[1:])
            F = int(f_ptr[-1])
            f_pos = np.repeat(a0 - f_ptr[:-1], n_f) + np.arange(F, dtype=np.int64)
            fvals = (ae[f_pos] - self.offset) if F else np.zeros(0, np.int64)
            self.filt_ptr, self.filt_idx = _sorted_unique_rows(np.repeat(np.arange(P), n_f), fvals, P)
            # all_splits lists are sets (unique) in the reference; their stored order defines the candidate
            # order of batch-shared evaluation batches (openkge/dataset.py:821-825)
            self.filt_idx_unsorted = fvals.astype(np.int32) if F == int(self.filt_ptr[-1]) else self.filt_idx
        else:
            self.filt_ptr, self.filt_idx = np.zeros(P + 1, np.int64), np.zeros(0, np.int32)
            self.filt_idx_unsorted = self.filt_idx

    @classmethod
    def from_csr(cls, prefix, slot, lab_ptr, lab_idx, n_cols: int, offset: int = 2, is_training_data: bool = True,
                 ans_ptr=None, alt_ptr=None, alt_idx=None, filt_ptr=None, filt_idx=None) -> "PrefixIndex":
        """Build from already-decoded CSR arrays (synthetic graphs, or a cache of a decoded split)."""
        self = cls.__new__(cls)
        P = len(prefix)
        self.n_cols, self.offset, self.is_training_data = int(n_cols), int(offset), is_training_data
        self.prefix = np.asarray(prefix, np.int32).reshape(-1, 2)
        self.slot = np.asarray(slot, np.int32)
        self.lab_ptr = np.asarray(lab_ptr, np.int64)
        self.lab_idx = np.asarray(lab_idx, np.int32)
        if ans_ptr is None:                       # one ranked answer per label, no alternatives
            ans_ptr = self.lab_ptr
            alt_ptr = np.arange(len(self.lab_idx) + 1, dtype=np.int64)
            alt_idx = self.lab_idx
        self.ans_ptr = np.asarray(ans_ptr, np.int64)
        self.alt_ptr = np.asarray(alt_ptr, np.int64)
        self.alt_idx = np.asarray(alt_idx, np.int32)
        self.filt_ptr = np.asarray(filt_ptr, np.int64) if filt_ptr is not None else np.zeros(P + 1, np.int64)
        self.filt_idx = np.asarray(filt_idx, np.int32) if filt_idx is not None else np.zeros(0, np.int32)
        self.filt_idx_unsorted = self.filt_idx
        return self

    def __len__(self) -> int:
        return len(self.prefix)

    def collate(self, rows: Sequence[int], pin: bool = False):
        """Batch of prefix rows in the reference's 7-tuple layout with sparse labels. Rows are regrouped
        po (slot 0) first, then sp (slot 2), each in batch order (openkge/dataset.py:794-811, 885-932)."""
        rows = np.asarray(rows, dtype=np.int64).reshape(-1)
        slot = self.slot[rows]
        order = np.concatenate([rows[slot == 0], rows[slot == 2]])
        n_po = int((slot == 0).sum())
        B = len(order)
        pref = self.prefix[order]

        def pair(block):
            if len(block) == 0:
                return None
            t = torch.from_numpy(np.ascontiguousarray(block))
            return (t[:, 0:1].contiguous(), t[:, 1:2].contiguous())       # .chunk(2, dim=1), :932

        slot_inputs = [pair(pref[:n_po]), pair(pref[n_po:])]
        lp, li = _csr_take(self.lab_ptr, self.lab_idx, order)
        labels = CSRMatrix(torch.from_numpy(lp.astype(np.int32)), torch.from_numpy(li.astype(np.int32)), (B, self.n_cols))
        normalizer_metric = float(len(li))                                  # label_tensor.sum(), :934
        normalizer_loss = B * self.n_cols                                   # :935
        shared = AllEntityIds(self.offset, self.n_cols)                      # arange(...)[offset:], :872
        label_ids = filt = None
        if not self.is_training_data:
            fp, fi = _csr_take(self.filt_ptr, self.filt_idx, order)
            filt = CSRMatrix(torch.from_numpy(fp.astype(np.int32)), torch.from_numpy(fi.astype(np.int32)), (B, self.n_cols))
            k = (self.ans_ptr[order + 1] - self.ans_ptr[order]).astype(np.int64)
            ans_row = np.repeat(np.arange(B, dtype=np.int32), k)
            a_ptr = np.zeros(B + 1, np.int64)
            np.cumsum(k, out=a_ptr[1:])
            # answers of the selected prefixes, then their alternative lists
            ans_ids = np.repeat(self.ans_ptr[order] - a_ptr[:-1], k) + np.arange(int(a_ptr[-1]), dtype=np.int64)
            ap, ai = _csr_take(self.alt_ptr, self.alt_idx, ans_ids)
            label_ids = RankedAnswers(torch.from_numpy(ans_row), torch.from_numpy(ap.astype(np.int32)),
                                      torch.from_numpy(ai.astype(np.int32)), n_rows=B)
        if pin:
            labels = labels.pin_memory()
            slot_inputs = [None if s is None else tuple(t.pin_memory() for t in s) for s in slot_inputs]
            if filt is not None:
                filt, label_ids = filt.pin_memory(), label_ids.pin_memory()
        return slot_inputs, normalizer_loss, normalizer_metric, labels, label_ids, filt, shared


def _pad4(n: int) -> int:
    return (int(n) + 3) // 4 * 4


class PackedBatch:
    """A training batch as ONE contiguous int32 array ``packed`` = [entity ids | relation ids | label row pointer, n_po |
    label columns] (every section padded to a multiple of 4 entries, ``packed_layout``) that behaves as the reference's
    7-tuple ``(slot_inputs, normalizer_loss, normalizer_metric, labels, None, None, batch_shared_entities)``: unpacking /
    indexing it builds the tuple's tensors as views into ``packed`` on first use. A graphed step never does: it copies
    ``packed`` to the device with a single H2D copy (``graphed.GraphedTrainStep.load``) and derives the po / sp row kinds
    and batch-norm segments from ``n_po`` (the number of po rows; they come first) on the device."""

    __slots__ = ("packed", "rows", "n_po", "nnz", "n_cols", "shared", "_items")

    def __init__(self, packed: torch.Tensor, rows: int, n_po: int, nnz: int, n_cols: int, shared):
        self.packed, self.rows, self.n_po, self.nnz, self.n_cols, self.shared = packed, int(rows), int(n_po), int(nnz), int(n_cols), shared
        self._items = None

    @property
    def normalizer_loss(self) -> int:
        return self.rows * self.n_cols

    @property
    def normalizer_metric(self) -> float:
        return float(self.nnz)

    def _tuple(self):
        if self._items is None:
            B, p = self.rows, self.n_po
            _, o_rel, o_ptr, o_idx = packed_layout(B)
            ent = self.packed[:B].unsqueeze(-1)                          # [B, 1]: the reference's id columns
            rel = self.packed[o_rel:o_rel + B].unsqueeze(-1)
            po = None if p == 0 else (rel[:p], ent[:p])                  # po prefix = (relation, object)
            sp = None if p == B else (ent[p:], rel[p:])                  # sp prefix = (subject, relation)
            labels = CSRMatrix(self.packed[o_ptr:o_ptr + B + 1], self.packed[o_idx:o_idx + self.nnz], (B, self.n_cols))
            self._items = ([po, sp], self.normalizer_loss, self.normalizer_metric, labels, None, None, self.shared)
        return self._items

    def __len__(self) -> int:
        return 7

    def __iter__(self):
        return iter(self._tuple())

    def __getitem__(self, i):
        if i == 1:
            return self.normalizer_loss
        if i == 2:
            return self.normalizer_metric
        return self._tuple()[i]

    @property
    def nbytes(self) -> int:
        return int(self.packed.numel()) * 4


def packed_layout(rows: int):
    """Offsets (in int32 entries) of the sections of ``PackedBatch.packed``: ent, rel, ptr (rows + 1 entries, followed by
    n_po), idx."""
    o_rel = _pad4(rows)
    o_ptr = o_rel + _pad4(rows)
    o_idx = o_ptr + _pad4(rows + 2)
    return 0, o_rel, o_ptr, o_idx


def collate_many(index: "PrefixIndex", rows_2d: np.ndarray, pin: bool = False):
    """``[index.collate(r) for r in rows_2d]`` for TRAINING batches of equal size, built by one native call for all k
    batches (``okge_host_collate_*``): one pinned allocation, the batches are views into it (``PackedBatch``). The
    per-batch interpreter time of the collate is what bounds the input rate of the small, graph-replayed steps: 0.26 ms
    per 512 rows for ``collate``, 0.09 ms for the numpy form of this function, against a 0.13 ms step at FB15k-237
    size."""
    if not index.is_training_data:
        raise ValueError("collate_many builds training batches (evaluation batches carry ragged answer / filter lists)")
    from . import _capi
    rows_2d = np.ascontiguousarray(rows_2d, dtype=np.int64)
    k, B = rows_2d.shape
    cache = index.__dict__.get("_row_columns")
    if cache is None:
        # per prefix row: is it an sp row, its entity id and its relation id. po rows are (relation, object), sp rows
        # (subject, relation): the entity id is column 1 of a po row and column 0 of an sp row
        sp_all = index.slot == 2
        cache = index.__dict__["_row_columns"] = (
            np.ascontiguousarray(sp_all, dtype=np.uint8),
            np.ascontiguousarray(np.where(sp_all, index.prefix[:, 0], index.prefix[:, 1]), dtype=np.int32),
            np.ascontiguousarray(np.where(sp_all, index.prefix[:, 1], index.prefix[:, 0]), dtype=np.int32),
            np.ascontiguousarray(index.lab_ptr, dtype=np.int64), np.ascontiguousarray(index.lab_idx, dtype=np.int32))
    sp_all, ent_of_row, rel_of_row, lab_ptr, lab_idx = cache
    # The work is done by two plain-C entry points of the library (csrc/host_collate.cu; the calls release the
    # interpreter lock, so a loader thread runs next to the thread that launches the GPU work): stable po | sp partition
    # of every batch, id columns, CSR labels, written into ONE (pinned) int32 buffer in which batch b owns
    # [starts[b], starts[b + 1]) in the PackedBatch layout [ent | rel | ptr, n_po | idx]. The tensors of the 7-tuples
    # are views into it.
    lib = _capi.load()
    _, o_rel, o_ptr, o_idx = packed_layout(B)
    counts, starts, n_po = np.empty(k, np.int64), np.empty(k + 1, np.int64), np.empty(k, np.int32)
    _capi.check(lib.okge_host_collate_plan(rows_2d.ctypes.data, k, B, lab_ptr.ctypes.data, len(sp_all), counts.ctypes.data,
                                           starts.ctypes.data), "okge_host_collate_plan")
    packed = torch.empty(int(starts[-1]), dtype=torch.int32, pin_memory=bool(pin))
    _capi.check(lib.okge_host_collate_fill(rows_2d.ctypes.data, k, B, sp_all.ctypes.data, ent_of_row.ctypes.data,
                                           rel_of_row.ctypes.data, lab_ptr.ctypes.data, lab_idx.ctypes.data, starts.ctypes.data,
                                           packed.data_ptr(), n_po.ctypes.data), "okge_host_collate_fill")
    shared = AllEntityIds(index.offset, index.n_cols)
    starts_l, n_po_l, counts_l = starts.tolist(), n_po.tolist(), counts.tolist()
    return [PackedBatch(packed[starts_l[b]:starts_l[b] + o_idx + counts_l[b]], B, n_po_l[b], counts_l[b], index.n_cols, shared)
            for b in range(k)]


def collate_shared(index: "PrefixIndex", rows: Sequence[int], min_size_batch_labels: int = 0, pin: bool = False,
                   exact_sampling: bool = True):
    """Batch-shared-entities mode of the reference collate (``use_batch_shared_entities=True``,
    openkge/dataset.py:813-868, 899-919): the candidates of a batch are the entities that occur as answers in
    the batch (training: this split's answers; eval: the all-splits filter sets), in first-occurrence order over
    the batch rows, topped up to ``min_size_batch_labels`` with uniformly sampled negatives; label / filter /
    answer columns are positions in that list. Sampling uses ``numpy.random.choice`` on the global numpy
    generator and a Python ``set`` exactly like the reference (:851-860), so the same ``numpy.random.seed``
    yields the same candidate list. ``exact_sampling=False`` draws the negatives without the full permutation of all
    entity ids that ``choice(replace=False)`` performs (50+ ms per batch at 2.5 M entities): the same distribution -
    a uniformly random set of distinct non-candidate ids - from a different random stream."""
    rows = np.asarray(rows, dtype=np.int64).reshape(-1)
    off = index.offset
    if index.is_training_data:
        _, seen = _csr_take(index.alt_ptr, index.alt_idx, _answers_of(index, rows))
    else:
        _, seen = _csr_take(index.filt_ptr, index.filt_idx_unsorted, rows)
    # ordered first occurrences (OrderedDict semantics, :813-825)
    _, first = np.unique(seen, return_index=True)
    keys = (seen[np.sort(first)].astype(np.int64) + off)
    min_size = 0 if min_size_batch_labels is None or min_size_batch_labels < 0 else int(min_size_batch_labels)
    # entity id -> local column: a persistent table (20 MB at 2.5 M entities), only the touched entries are reset afterwards
    lut = index.__dict__.get("_shared_lut")
    if lut is None or len(lut) != index.n_cols + off:
        lut = np.full(index.n_cols + off, -1, dtype=np.int64)
        index.__dict__["_shared_lut"] = lut
    if len(keys) >= min_size:
        shared = keys
    elif exact_sampling:
        negatives = set((np.random.choice(index.n_cols, min_size, replace=False) + off).tolist())   # :851-853
        negatives.difference_update(keys.tolist())
        shared = np.asarray((keys.tolist() + list(negatives))[:min_size], dtype=np.int64)            # :855-858
    else:
        need = min_size - len(keys)
        lut[keys] = 0                                                             # mark the candidates we already have
        draw = np.random.randint(0, index.n_cols, size=2 * need + 64) + off       # with replacement, then distinct
        _, first_neg = np.unique(draw, return_index=True)
        cand = draw[np.sort(first_neg)]                                           # in order of arrival
        cand = cand[lut[cand] < 0][:need]
        if len(cand) < need:                                                      # (practically unreachable)
            rest = np.setdiff1d(np.arange(off, index.n_cols + off), np.concatenate([keys, cand]))
            cand = np.concatenate([cand, np.random.permutation(rest)[:need - len(cand)]])
        shared = np.concatenate([keys, cand]).astype(np.int64)
    n_local = len(shared)
    lut[shared] = np.arange(n_local)
    slot = index.slot[rows]
    order = np.concatenate([rows[slot == 0], rows[slot == 2]])
    n_po = int((slot == 0).sum())
    B = len(order)
    pref = index.prefix[order]

    def pair(block):
        if len(block) == 0:
            return None
        t = torch.from_numpy(np.ascontiguousarray(block))
        return (t[:, 0:1].contiguous(), t[:, 1:2].contiguous())

    def remap(ptr, idx):
        loc = lut[idx.astype(np.int64) + off]
        row_of = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
        p2, i2 = _sorted_unique_rows(row_of, loc, len(ptr) - 1)
        return p2, i2

    lp, li = _csr_take(index.lab_ptr, index.lab_idx, order)
    lp, li = remap(lp, li)
    labels = CSRMatrix(torch.from_numpy(lp.astype(np.int32)), torch.from_numpy(li.astype(np.int32)), (B, n_local))
    label_ids = filt = None
    if not index.is_training_data:
        fp, fi = _csr_take(index.filt_ptr, index.filt_idx, order)
        fp, fi = remap(fp, fi)
        filt = CSRMatrix(torch.from_numpy(fp.astype(np.int32)), torch.from_numpy(fi.astype(np.int32)), (B, n_local))
        ans_ids = _answers_of(index, order)
        k = (index.ans_ptr[order + 1] - index.ans_ptr[order]).astype(np.int64)
        ap, ai = _csr_take(index.alt_ptr, index.alt_idx, ans_ids)
        label_ids = RankedAnswers(torch.from_numpy(np.repeat(np.arange(B, dtype=np.int32), k)),
                                  torch.from_numpy(ap.astype(np.int32)),
                                  torch.from_numpy(lut[ai.astype(np.int64) + off].astype(np.int32)), n_rows=B)
    lut[shared] = -1                                                     # hand the table back clean
    shared_t = torch.from_numpy(shared.astype(np.int32)).unsqueeze(1)
    slot_inputs = [pair(pref[:n_po]), pair(pref[n_po:])]
    if pin:
        labels, shared_t = labels.pin_memory(), shared_t.pin_memory()
        slot_inputs = [None if s is None else tuple(t.pin_memory() for t in s) for s in slot_inputs]
        if filt is not None:
            filt, label_ids = filt.pin_memory(), label_ids.pin_memory()
    return slot_inputs, B * n_local, float(len(li)), labels, label_ids, filt, shared_t


class DeviceRows:
    """One training batch named by its prefix-row indices only (int64 DEVICE tensor [B]): what a loader yields when the
    collate itself runs on the GPU (``DeviceSharedCollate``). No host work and no H2D copy per step."""

    def __init__(self, rows: torch.Tensor):
        self.rows = rows

    def __len__(self) -> int:
        return int(self.rows.numel())


class DeviceSharedCollate:
    """Training-mode batch-shared collate of the reference (``use_batch_shared_entities=True``, openkge/dataset.py:813-868,
    899-919) ON THE DEVICE (``okge_collate_shared``, csrc/collate_ops.cu): from B prefix-row indices to the tensors of the
    training step — entity / relation ids with the po rows first, CSR labels whose columns are positions in the batch's
    candidate list, the candidate list itself (the entities that occur as answers in the batch, topped up to ``min_size``
    with uniformly drawn distinct negatives), its length and 1 / (B * length). Every buffer has a FIXED shape (capacities
    instead of data-dependent sizes), nothing synchronises with the host, and the six launches can be captured in the
    same CUDA graph as the training step they feed.

    Same distribution as the host collate, not the same stream: candidates are ordered by entity id instead of first
    occurrence (the order of the columns does not enter the loss) and the negatives come from Philox keyed by
    (seed, call number) (``dataset.collate_shared`` stays the reference-exact path: same ``numpy.random.choice`` stream,
    same order)."""

    def __init__(self, index: "PrefixIndex", min_size: int, cap_nnz: int, cap_cols: int, device, seed: int = 0,
                 rows_per_batch: Optional[int] = None):
        if not index.is_training_data:
            raise ValueError("the device collate builds training batches (evaluation keeps the host collate)")
        self.min_size = max(int(min_size), 0)
        self.cap_nnz, self.cap_cols = int(cap_nnz), int(cap_cols)
        self.n_entities, self.offset, self.seed = int(index.n_cols), int(index.offset), int(seed)
        self.device = torch.device(device)
        lab_ptr_h, lab_idx_h = np.asarray(index.lab_ptr, dtype=np.int64), np.asarray(index.lab_idx)
        inside = np.ones(len(lab_idx_h), bool)                     # label j is not the first of its row
        inside[lab_ptr_h[:-1][lab_ptr_h[:-1] < len(lab_idx_h)]] = False
        if len(lab_idx_h) > 1 and bool((inside[1:] & (lab_idx_h[1:] <= lab_idx_h[:-1])).any()):
            raise ValueError("the device collate needs every answer list of the index ascending and duplicate-free")
        dev = self.device
        self.lab_ptr = torch.from_numpy(np.ascontiguousarray(index.lab_ptr, dtype=np.int64)).to(dev)      # [P + 1]
        self.lab_idx = torch.from_numpy(np.ascontiguousarray(index.lab_idx, dtype=np.int32)).to(dev)      # entity id - offset
        self.prefix = torch.from_numpy(np.ascontiguousarray(index.prefix, dtype=np.int32)).to(dev)        # [P, 2]
        self.slot = torch.from_numpy(np.ascontiguousarray(index.slot, dtype=np.int32)).to(dev)            # 0 = po, 2 = sp
        self.n_draw = 2 * self.min_size + 64
        n_words = (self.n_entities + 31) // 32
        i32 = dict(dtype=torch.int32, device=dev)
        self.ws = dict(bitmap=torch.zeros(n_words, **i32), word_prefix=torch.zeros(n_words, **i32),
                       tile_sum=torch.zeros((n_words + 1023) // 1024, **i32),
                       first_draw=torch.full((self.n_entities,), 2 ** 31 - 1, **i32),
                       e_flat=torch.zeros(self.cap_nnz + self.n_draw, **i32), row_start=None)
        self.scalars = torch.zeros(8, dtype=torch.int64, device=dev)
        self._out = None
        if rows_per_batch is not None:
            self._buffers(int(rows_per_batch))

    # accumulated over the calls (0-dim views of the scalars): batches cut to a capacity, positive labels handed out
    @property
    def overflow(self) -> torch.Tensor:
        return self.scalars[4]

    @property
    def nnz_total(self) -> torch.Tensor:
        return self.scalars[5]

    def _buffers(self, B: int) -> dict:
        if self._out is None or self._out["ent"].numel() != B:
            i32 = dict(dtype=torch.int32, device=self.device)
            self.ws["row_start"] = torch.zeros(B, dtype=torch.int64, device=self.device)
            self._out = dict(ent=torch.zeros(B, **i32), rel=torch.zeros(B, **i32), is_po=torch.zeros(B, **i32),
                             ptr=torch.zeros(B + 1, **i32), idx=torch.full((self.cap_nnz,), -1, **i32),
                             cand=torch.zeros(self.cap_cols, **i32), scalars=self.scalars,
                             count=torch.zeros(1, **i32), inv_norm=torch.zeros(1, dtype=torch.float32, device=self.device))
        return self._out

    def __call__(self, rows: torch.Tensor, out: Optional[dict] = None):
        """Returns a dict of device tensors (the collate's own buffers, overwritten by the next call, unless ``out`` names
        others): ent, rel [B, 1] int32; is_po [B] int32; ptr [B + 1] int32; idx [cap_nnz] int32 (-1 behind nnz); cand
        [cap_cols, 1] int32 (entity ids; rows behind count repeat a valid id); count [1] int32; inv_norm [1] fp32 =
        1 / (B * count); b_po, nnz: 0-dim int64 views of the scalars."""
        from . import kernels as K
        buf = dict(self._buffers(rows.numel()))
        if out:
            buf.update(out)
        K.collate_shared(rows, self.lab_ptr, self.lab_idx, self.prefix, self.slot, self.n_entities, self.offset, self.min_size,
                         self.cap_nnz, self.cap_cols, self.n_draw, self.seed, self.ws, buf)
        return dict(ent=buf["ent"].view(-1, 1), rel=buf["rel"].view(-1, 1), is_po=buf["is_po"], ptr=buf["ptr"], idx=buf["idx"],
                    cand=buf["cand"].view(-1, 1), count=buf["count"], inv_norm=buf["inv_norm"],
                    b_po=self.scalars[0], nnz=self.scalars[2])


def _answers_of(index: "PrefixIndex", rows: np.ndarray) -> np.ndarray:
    """Flat ids of the ranked answers of the given prefix rows, in row order."""
    k = (index.ans_ptr[rows + 1] - index.ans_ptr[rows]).astype(np.int64)
    a_ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(k, out=a_ptr[1:])
    return np.repeat(index.ans_ptr[rows] - a_ptr[:-1], k) + np.arange(int(a_ptr[-1]), dtype=np.int64)


def input_and_labels_to_device(data, training: bool, device, non_blocking: bool = True):
    """``OneToNMentionRelationDataset.input_and_labels_to_device`` (openkge/dataset.py:385-421) for the
    sparse batch: moves prefixes, labels, (eval) filters and answer lists; the 1-vs-all
    ``batch_shared_entities`` stay symbolic (an arange) instead of an [N, 1] H2D copy."""
    slot_inputs, nl, nm, labels, label_ids, filt, shared = data
    moved = []
    for s in slot_inputs:
        moved.append(None if s is None else tuple(t.to(device, non_blocking=non_blocking) for t in s))
    labels = labels.to(device, non_blocking=non_blocking) if isinstance(labels, (CSRMatrix, torch.Tensor)) else labels
    if not training:
        if filt is not None:
            filt = filt.to(device, non_blocking=non_blocking)
        if isinstance(label_ids, RankedAnswers):
            label_ids = label_ids.to(device, non_blocking=non_blocking)
    if shared is not None:
        shared = shared.to(device, non_blocking=non_blocking)
    return moved, nl, nm, labels, label_ids, filt, shared


def batch_h2d_bytes(data) -> int:
    if isinstance(data, PackedBatch):
        return data.nbytes
    slot_inputs, _, _, labels, label_ids, filt, shared = data
    n = sum(t.numel() * t.element_size() for s in slot_inputs if s is not None for t in s)
    if isinstance(shared, torch.Tensor):
        n += shared.numel() * shared.element_size()
    n += labels.nbytes if isinstance(labels, CSRMatrix) else labels.numel() * labels.element_size()
    if filt is not None:
        n += filt.nbytes if isinstance(filt, CSRMatrix) else filt.numel() * filt.element_size()
    if isinstance(label_ids, RankedAnswers):
        n += label_ids.nbytes
    return int(n)


# ---------------------------------------------------------------------------------------------
# filtered ranking — drop-in for OneToNMentionRelationDataset.compute_metrics
# ---------------------------------------------------------------------------------------------

def _as_csr(x, device) -> CSRMatrix:
    if isinstance(x, CSRMatrix):
        return x.to(device)
    return CSRMatrix.from_dense(x.to(device))


def _as_answers(label_ids, device) -> RankedAnswers:
    if isinstance(label_ids, RankedAnswers):
        return label_ids.to(device)
    return RankedAnswers.from_label_ids(label_ids).to(device)


def rank_answers(filter_mask, label_ids, predictions) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, RankedAnswers]:
    """(true_score, greater, equal, answers) for every ranked answer of the batch.

    Dense ``predictions`` [B, N]: one pass of ``okge_rank_count`` over the materialised matrix.
    ``PrefixScores`` (Q, E): fused path — scores of the few label / filter columns come from the same
    tensor-core kernel on the gathered candidate rows, the dense count-greater runs inside the scoring
    epilogue over all N candidates, then the filtered columns are subtracted (integer arithmetic)."""
    if isinstance(predictions, PrefixScores):
        dev = predictions.q.device
        filt, ans = _as_csr(filter_mask, dev), _as_answers(label_ids, dev)
        if predictions.shard is not None:            # candidates partitioned over ranks: integer counts summed over ranks
            from .sharded import sharded_rank_counts
            sh = predictions.shard
            q16, e16 = predictions.operands()
            predictions.ensure_loss()
            return (*sharded_rank_counts(K, sh.comm, predictions.q, predictions.e, sh.lo, sh.hi, sh.rank, ans, filt, e16=e16,
                                         split=predictions.split), ans)
        return (*_rank_fused(predictions, filt, ans), ans)
    dev = predictions.device
    if not predictions.is_cuda:
        raise RuntimeError("compute_metrics needs CUDA predictions; the B200 path has no CPU fallback")
    filt, ans = _as_csr(filter_mask, dev), _as_answers(label_ids, dev)
    true, greater, equal = K.rank_count(predictions, ans.ans_row, ans.alt_ptr, ans.alt_idx, filt.ptr, filt.idx)
    return true, greater, equal, ans


def _rank_fused(ps: PrefixScores, filt: CSRMatrix, ans: RankedAnswers):
    dev = ps.q.device
    Q = len(ans)
    greater = torch.zeros(Q, dtype=torch.int32, device=dev)
    equal = torch.zeros(Q, dtype=torch.int32, device=dev)
    true = torch.full((Q,), float("-inf"), dtype=torch.float32, device=dev)
    if Q == 0:
        return true, greater, equal
    # 1) scores of the sparse columns (answers' alternatives + filters) from the SAME GEMM arithmetic. Every listed
    #    column is gathered as it is (duplicates included): no unique / sort, hence no host synchronisation, and the
    #    small [B, n_alt + n_filt] product costs microseconds next to the pass over all N candidates.
    cols = torch.cat([ans.alt_idx, filt.idx]).to(torch.int32)
    n_alt = ans.alt_idx.numel()
    q16, e16 = ps.operands()
    split = ps.split
    sel = K.score_store(q16, K.gather_rows_f16(e16, cols), split=split)      # [B, n_alt + n_filt]
    pos = torch.arange(cols.numel(), dtype=torch.int32, device=dev)
    alt_pos, filt_pos = pos[:n_alt], pos[n_alt:]
    K.rank_true_score(sel, ans.ans_row, ans.alt_ptr, alt_pos, true)
    pend = ps.pending_loss
    if pend is not None and ans.overflow is not None:
        # 2a) evaluation step in ONE pass over the candidates: the loss sum and the counts of every ranked answer. A prefix
        #     row holds SLOTS answers; the (host-known, usually few) further answers ride on extra query rows behind the B
        #     prefix rows, which are scored and counted like the others but carry no loss
        B, S = ps.q.size(0), ans.slots
        n_x = int(ans.extra_prefix.numel())
        row = ans.ans_row.long()
        slot = torch.arange(Q, device=dev) - torch.searchsorted(ans.ans_row, ans.ans_row)    # position inside its row
        flat = torch.where(slot < S, row * 4 + slot, torch.zeros_like(row))                  # [B + n_x, 4] layout
        if n_x:
            flat[ans.overflow] = B * 4 + ans.overflow_slot
            q_all = K.gather_rows_f16(q16, torch.cat([torch.arange(B, dtype=torch.int32, device=dev),
                                                      ans.extra_prefix.to(torch.int32)]))
        else:
            q_all = q16
        thresh = torch.full(((B + n_x) * 4,), float("inf"), dtype=torch.float32, device=dev)
        thresh[flat] = true
        g4 = torch.zeros((B + n_x) * 4, dtype=torch.int32, device=dev)
        e4 = torch.zeros((B + n_x) * 4, dtype=torch.int32, device=dev)
        K.score_bce_rank(q_all, e16, pend["ptr"], pend["idx"], pend["y_base"], pend["y_pos"], thresh, g4, e4, pend["out"],
                         extra_rows=n_x, split=split, slots=S)
        ps.pending_loss = None
        greater.copy_(g4[flat])
        equal.copy_(e4[flat])
    else:
        # 2b) dense count over all candidates inside the scoring epilogue, one query row per ranked answer
        q_exp = K.gather_rows_f16(q16, ans.ans_row)
        K.score_rank(q_exp, e16, true, greater, equal, split=split)
    # 3) remove the filtered columns (and account for the -1e8 fill value itself)
    K.rank_filter_correct(sel, ans.ans_row, filt.ptr, filt_pos, true, greater, equal, add_mask_terms=True)
    return true, greater, equal


def metric_sums(greater: torch.Tensor, equal: torch.Tensor) -> torch.Tensor:
    """Device tensor of the six sums behind the rank meters: rank = greater + equal // 2 (openkge/dataset.py:445),
    then sum 1/(rank+1), sum rank, #rank<50, <10, <3, <1 (:446-452)."""
    ranks = greater.long() + torch.div(equal.long(), 2, rounding_mode="floor")
    return torch.stack([
        (1.0 / (ranks + 1).double()).sum(),
        ranks.double().sum(),
        (ranks < 50).double().sum(),
        (ranks < 10).double().sum(),
        (ranks < 3).double().sum(),
        (ranks < 1).double().sum(),
    ])


def metrics_from_sums(sums: Sequence[float], Q: int) -> MetricResult:
    result = MetricResult()
    if Q:
        for key, s in zip(("mrr", "mr", "h50", "h10", "h3", "h1"), sums):
            result[key].update(s / Q, Q)
    return result


def metrics_from_counts(greater: torch.Tensor, equal: torch.Tensor) -> MetricResult:
    """The six rank meters of the reference (openkge/dataset.py:445-452). Every meter is a per-prefix mean weighted by
    the prefix's answer count, i.e. the mean over all ranked answers; one D2H read of six sums."""
    Q = int(greater.numel())
    if Q == 0:
        return MetricResult()
    return metrics_from_sums(metric_sums(greater, equal).cpu().tolist(), Q)


def compute_metrics(filter_mask, label_ids, predictions) -> MetricResult:
    """Drop-in for ``OneToNMentionRelationDataset.compute_metrics`` (openkge/dataset.py:423-453).
    Accepts the reference's dense bool mask / list-of-lists label ids / dense predictions as well as
    the sparse wire types."""
    _, greater, equal, _ = rank_answers(filter_mask, label_ids, predictions)
    return metrics_from_counts(greater, equal)


class OneToNMentionRelationDataset:
    """Holder of one split in the B200 path: the decoded :class:`PrefixIndex` plus the loop-facing
    attributes ``Trainer`` reads (device, batch_size, input_style, use_batch_shared_entities)."""

    input_style = "right_and_left_prefix"
    use_batch_shared_entities = False
    batch_size_for_backward = None

    def __init__(self, index: PrefixIndex, meta: EntityRelationDatasetMeta, batch_size: int, device="cuda",
                 is_training_data: bool = True, use_batch_shared_entities: bool = False,
                 min_size_batch_labels: int = -1, exact_negative_sampling: bool = True):
        self.index, self.meta, self.batch_size, self.device = index, meta, batch_size, device
        self.is_training_data = is_training_data
        # True: the negatives of a batch-shared candidate list come from the reference's own random stream
        # (numpy.random.choice without replacement = a full permutation of all entity ids per batch); False: same
        # distribution without the permutation (not a reference option)
        self.exact_negative_sampling = bool(exact_negative_sampling)
        # train_data_config / val_data_config keys of the reference (openkge/default.yaml:121-156)
        self.use_batch_shared_entities = bool(use_batch_shared_entities)
        self.min_size_batch_labels = min_size_batch_labels

    def collate(self, rows, pin: bool = False):
        if self.use_batch_shared_entities:
            return collate_shared(self.index, rows, self.min_size_batch_labels, pin=pin,
                                  exact_sampling=getattr(self, "exact_negative_sampling", True))
        return self.index.collate(rows, pin=pin)

    def __len__(self) -> int:
        return len(self.index)

    def get_dataset_meta_dict(self) -> EntityRelationDatasetMeta:
        return self.meta

    def input_and_labels_to_device(self, data, training, device):
        return input_and_labels_to_device(data, training, device)

    compute_metrics = staticmethod(compute_metrics)

    def get_row_loader(self, shuffle: bool = True, seed: int = 0) -> "DeviceRowLoader":
        """Training batches as device row indices (see ``DeviceRowLoader``); ``Trainer.train_epoch`` collates them on the
        GPU inside the CUDA graph of the step (batch-shared candidate mode)."""
        return DeviceRowLoader(self, seed=seed, shuffle=shuffle)

    def get_loader(self, shuffle: bool = False, sampler: Optional[Sequence[int]] = None, drop_last: bool = True,
                   pin_memory: bool = True, seed: int = 0, prefetch: int = 0, reshuffle: bool = True):
        """Iterable of collated batches (the reference uses a torch DataLoader with forked workers,
        openkge/dataset.py:455-479; the vectorised collate does not need them). ``shuffle``: a fresh permutation every
        time the loader is iterated, like ``DataLoader(shuffle=True)`` (epoch e draws it from ``default_rng(seed + e)``;
        ``reshuffle=False`` keeps the permutation of ``seed`` for every epoch, for tests). ``prefetch`` > 0: a background
        thread collates (and pins) up to that many chunks ahead while the caller queues GPU work; same batches, same
        order."""
        n = len(self.index)
        order = None
        if sampler is not None:
            order = np.asarray(list(sampler), dtype=np.int64)
        elif not shuffle:
            order = np.arange(n)
        bs = self.batch_size
        it = _BatchIter(self, order, n, bs, drop_last, pin_memory and torch.cuda.is_available(), seed, reshuffle)
        return _Prefetcher(it, prefetch) if prefetch > 0 else it


class DeviceRowLoader:
    """Epochs of ``DeviceRows`` batches: the shuffled order of the prefix rows is drawn ON the device (one ``randperm`` per
    epoch) and a batch is a slice of it, for training loops whose collate runs on the GPU (``DeviceSharedCollate`` inside
    ``graphed.GraphedTrainStep``). Ragged last batches are dropped (``drop_last`` of the reference's training loader)."""

    def __init__(self, dataset, seed: int = 0, shuffle: bool = True):
        self.dataset, self.seed, self.shuffle, self.epoch = dataset, int(seed), bool(shuffle), 0

    def __len__(self):
        return len(self.dataset.index) // self.dataset.batch_size

    def __iter__(self):
        n, bs, dev = len(self.dataset.index), self.dataset.batch_size, self.dataset.device
        if self.shuffle:
            g = torch.Generator(device=dev)
            g.manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, device=dev, generator=g)
        else:
            order = torch.arange(n, device=dev)
        self.epoch += 1
        for i in range(0, (n // bs) * bs, bs):
            yield DeviceRows(order[i:i + bs])


class LazyPermutation:
    """A pseudo-random permutation of range(n) that is evaluated slice by slice instead of materialised: a 6-round Feistel
    network over the smallest even-width bit space that holds n, with cycle walking for the values that fall outside
    [0, n). ``DataLoader(shuffle=True)`` draws ``randperm(n)`` when an epoch starts (0.3 s of host time for the 15 M prefix
    rows of the 1 M-entity workload, before the first batch exists); here the epoch starts at once and a batch of B rows
    costs O(B). Bijective by construction (every round is invertible), keyed by ``seed``."""

    ROUNDS = 6

    def __init__(self, n: int, seed: int):
        self.n = int(n)
        self.half = max(((max(self.n, 2) - 1).bit_length() + 1) // 2, 1)
        self.mask = np.uint64((1 << self.half) - 1)
        self.keys = np.random.default_rng(seed).integers(1, 2 ** 62, self.ROUNDS).astype(np.uint64)

    def __len__(self) -> int:
        return self.n

    def _encrypt(self, x: np.ndarray) -> np.ndarray:
        h, mask = np.uint64(self.half), self.mask
        left, right = x >> h, x & mask
        for k in self.keys:
            f = (right + k) * np.uint64(0x9E3779B97F4A7C15)            # wraps mod 2^64
            f ^= f >> np.uint64(29)
            f *= np.uint64(0xBF58476D1CE4E5B9)
            f ^= f >> np.uint64(32)
            left, right = right, left ^ (f & mask)
        return (left << h) | right

    def __getitem__(self, key) -> np.ndarray:
        if not isinstance(key, slice):
            raise TypeError("slices only")
        start, stop, step = key.indices(self.n)
        y = self._encrypt(np.arange(start, stop, step, dtype=np.uint64))
        bad = np.flatnonzero(y >= np.uint64(self.n))
        while bad.size:                                             # cycle walking: the domain is < 4 n wide
            y[bad] = self._encrypt(y[bad])
            bad = bad[y[bad] >= np.uint64(self.n)]
        return y.astype(np.int64)


class _BatchIter:
    def __init__(self, dataset, order, n, bs, drop_last, pin, seed=0, reshuffle=True):
        self.index, self.order, self.n, self.bs, self.drop_last, self.pin = dataset, order, n, bs, drop_last, pin
        self.seed, self.reshuffle, self.epoch = int(seed), bool(reshuffle), 0
        m = n if order is None else len(order)
        self.stop = (m // bs) * bs if drop_last else m

    def __len__(self):
        return (self.stop + self.bs - 1) // self.bs

    CHUNK = 16          # full training batches collated per vectorised pass (collate_many)

    def __iter__(self):
        ds = self.index
        order = self.order
        if order is None:        # shuffle: a new permutation per epoch (a fixed one would drop the same tail every epoch)
            order = LazyPermutation(self.n, self.seed + (self.epoch if self.reshuffle else 0))
            self.epoch += 1
        many = (getattr(ds, "is_training_data", False) and not getattr(ds, "use_batch_shared_entities", False)
                and hasattr(ds, "index"))
        i = 0
        while i < self.stop:
            full = (self.stop - i) // self.bs
            if many and full >= 2:
                k = min(full, self.CHUNK)
                yield from collate_many(ds.index, order[i:i + k * self.bs].reshape(k, self.bs), pin=self.pin)
                i += k * self.bs
            else:
                yield ds.collate(order[i:min(i + self.bs, self.stop)], pin=self.pin)
                i += self.bs


class _Prefetcher:
    """Runs an iterable in a daemon thread, ``depth`` items ahead (numpy's gathers and the pinned copies release the GIL).
    A consumer that stops early (``break``, an exception, ``max_steps``) closes the generator, which tells the worker to
    stop: no thread and no pinned batches outlive the loop."""

    def __init__(self, inner, depth: int):
        self.inner, self.depth = inner, int(depth)

    def __len__(self):
        return len(self.inner)

    def __iter__(self):
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        done = object()
        stop = threading.Event()

        def put(item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.05)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for item in self.inner:
                    if not put(item):
                        return
                put(done)
            except BaseException as ex:  # noqa: BLE001  (re-raised in the consumer)
                put(ex)

        worker = threading.Thread(target=work, daemon=True)
        worker.start()
        try:
            while True:
                item = q.get()
                if item is done:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while True:                       # release whatever the worker has already queued
                try:
                    q.get_nowait()
                except queue.Empty:
                    break
            worker.join(timeout=1.0)
