"""Dataset cache build: the reference's on-disk id files -> the int32 tensors its collate consumes.

Drop-in for ``OneToNMentionRelationDataset._collect_seen_triples`` / ``merge_all_splits_triples`` /
``create_data_tensors`` / ``load_vocab`` (openkge/dataset.py:141-309, 481-710). The reference walks every line in
Python, round-trips the prefix groups through ``*.jsonl`` files and fills the tensors element by element (README: "around
30 minutes and up to 10-20 GB" on OLPBench). Here the same tensors come from a handful of vectorised numpy passes:
one stable string sort per direction, a boundary scan, and prefix sums for the packed ragged layout — a sort / segment job.

Outputs are bit-identical to the reference's (``tests/test_host_logic.py`` checks them against tensors dumped from the
unmodified reference), including its observable quirks, which the collate indices depend on:

* lines are ordered by the STRING form of the ids (``sorted(..., key=lambda l: l.split("\\t")[i])``, :495-500), i.e.
  "10" < "2"; the order inside a prefix group is the file order (stable sorts);
* the LAST prefix group of every direction of every split is never written (:489-518 has no flush after the loop), so it
  is missing from the split and from the all-splits filter lists (``drop_last_group=True`` reproduces this);
* all-splits answer sets are Python ``set`` objects dumped with ``list(...)`` (:544-560): their order is CPython's hash
  order. ``exact_set_order=True`` rebuilds them through real Python sets (bit-identical, one small Python loop per
  prefix); the default stores them sorted, which only changes the candidate ORDER of batch-shared evaluation batches;
* with ``max_size_prefix_label > 1`` long training groups are split into chunks of that many answers (:668-690). The
  reference also allocates one uninitialised phantom row per chunked group (:636-640 vs :673); those rows are not
  reproduced.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .dataset import EntityRelationDatasetMeta, PrefixIndex

# (name, relation column, entity column of the prefix, prefix = (col a, col b), slot id, column of the answer ids)
_DIRECTIONS = (("sp_o", 1, 0, (0, 1), 2, 4), ("po_s", 1, 2, (1, 2), 0, 3))


@dataclass
class Split:
    """One ``{train,valid,test}.txt`` file: 5 TAB columns ``s  r  o  s_alternatives  o_alternatives`` (openkge/default.yaml:
    104-116), alternatives space separated."""
    cols: List                        # per id column: ordering keys that sort like the reference's STRING comparison ...
    ids: np.ndarray                   # ... and the ids as int64 [n, 3]
    alt_ptr: Dict[int, np.ndarray]    # column 3 / 4 -> CSR pointer over lines
    alt_val: Dict[int, np.ndarray]    # column 3 / 4 -> flat alternative ids


_split_cache: Dict[Tuple[str, float, int], Split] = {}


def read_split(path: str) -> Split:
    """Parsed once per (path, mtime, size): building train / valid / test tensors reads every file three times."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime, st.st_size)
    if key not in _split_cache:
        if len(_split_cache) > 8:
            _split_cache.clear()
        _split_cache[key] = _parse_split(path)
    return _split_cache[key]


def _string_order_keys(v: np.ndarray) -> Tuple[np.ndarray, ...]:
    """Keys whose lexsort order equals Python's string order of the canonical decimal strings of the non-negative ints
    ``v`` (the reference sorts the id columns as text, openkge/dataset.py:489-492): digits left-aligned to a common width,
    the shorter string first when one is a prefix of the other. Returned minor key first, like np.lexsort wants them."""
    nd = np.ones(v.shape, np.int64)
    bound = 10
    while True:                                       # number of decimal digits (<= 18 iterations)
        more = v >= bound
        if not more.any():
            break
        nd += more
        bound *= 10
    width = int(nd.max()) if v.size else 1
    return nd, v * (10 ** (width - nd))


def _parse_split_text(path: str) -> Split:
    """Line-by-line parser: the fallback for files the vectorised parser does not accept (blank lines, non-canonical
    integers); the ordering keys are the strings themselves."""
    with open(path) as f:
        lines = [ln.rstrip("\n").split("\t") for ln in f if ln.strip()]
    n = len(lines)
    cols = [(np.array([ln[c] for ln in lines], dtype=np.str_),) for c in range(3)]
    ids = np.array([[int(ln[0]), int(ln[1]), int(ln[2])] for ln in lines], dtype=np.int64).reshape(n, 3)
    alt_ptr, alt_val = {}, {}
    for c in (3, 4):
        lists = [ln[c].split() for ln in lines]
        ptr = np.zeros(n + 1, np.int64)
        np.cumsum([len(x) for x in lists], out=ptr[1:])
        alt_ptr[c] = ptr
        alt_val[c] = np.array([int(v) for x in lists for v in x], dtype=np.int64)
    return Split(cols, ids, alt_ptr, alt_val)


def _parse_split(path: str) -> Split:
    """Vectorised parser of a 5-column split file: the bytes are scanned with numpy (delimiter positions, token counts per
    field), every integer of the file is converted in one C call, and the columns are index arithmetic on that flat
    array - no Python object per line (the cache build is parse-bound; OLPBench has 30 M lines)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size == 0:
        return _parse_split_text(path)
    if raw[-1] != 10:
        raw = np.append(raw, np.uint8(10))
    nl = np.flatnonzero(raw == 10)
    tab = np.flatnonzero(raw == 9)
    n = nl.size
    if tab.size != 4 * n or (raw == 13).any():
        return _parse_split_text(path)
    tabs = tab.reshape(n, 4)
    starts = np.concatenate([[0], nl[:-1] + 1])
    if not ((tabs[:, 0] > starts).all() and (tabs[:, 3] < nl).all()):
        return _parse_split_text(path)                                  # some line does not have its 4 tabs (or is blank)
    lo = np.stack([starts, tabs[:, 0] + 1, tabs[:, 1] + 1, tabs[:, 2] + 1, tabs[:, 3] + 1], 1)
    hi = np.stack([tabs[:, 0], tabs[:, 1], tabs[:, 2], tabs[:, 3], nl], 1)
    is_tok = (raw != 32) & (raw != 9) & (raw != 10)
    if not (((raw >= 48) & (raw <= 57)) | ~is_tok).all():
        return _parse_split_text(path)                                  # signs, letters, ...: let Python's int() decide
    tok_start = is_tok.copy()
    tok_start[1:] &= ~is_tok[:-1]
    cs = np.concatenate([[0], np.cumsum(tok_start)])
    cnt = cs[hi] - cs[lo]                                               # integers per field, [n, 5]
    if not (cnt[:, :3] == 1).all():
        return _parse_split_text(path)
    txt = raw.copy()
    txt[~is_tok] = 32
    flat = np.array(txt.tobytes().split(), dtype=np.int64)              # every integer of the file, in file order
    per_line = cnt.sum(1)
    line_off = np.concatenate([[0], np.cumsum(per_line)[:-1]])
    ids = np.stack([flat[line_off], flat[line_off + 1], flat[line_off + 2]], 1)
    nd_all = [hi[:, c] - lo[:, c] for c in range(3)]
    cols = []
    for c in range(3):
        nd, key = _string_order_keys(ids[:, c])
        if not (nd == nd_all[c]).all():
            return _parse_split_text(path)                              # leading zeros / padding: not canonical strings
        cols.append((nd, key))
    alt_ptr, alt_val = {}, {}
    first = line_off + 3
    for c in (3, 4):
        k = cnt[:, c]
        ptr = np.zeros(n + 1, np.int64)
        np.cumsum(k, out=ptr[1:])
        idx = np.repeat(first - ptr[:-1], k) + np.arange(int(ptr[-1]), dtype=np.int64)
        alt_ptr[c], alt_val[c] = ptr, flat[idx]
        first = first + k
    return Split(cols, ids, alt_ptr, alt_val)


@dataclass
class Groups:
    """Prefix groups of one direction of one split, in the reference's jsonl order."""
    prefix: np.ndarray      # [G, 2]
    slot: int
    line_ptr: np.ndarray    # [G + 1] answers (lines) per group
    alt_ptr: np.ndarray     # [#answers + 1] alternative mentions per answer
    alt_val: np.ndarray     # flat mention ids


def group_direction(split: Split, direction, drop_last_group: bool = True) -> Groups:
    """openkge/dataset.py:489-518 for one direction: stable sort by (entity string, relation string), consecutive equal
    prefixes form a group, the last group is lost."""
    _, rel_col, ent_col, (pa, pb), slot, ans_col = direction
    n = len(split.ids)
    # sorted(sorted(lines, key=rel), key=ent), both stable == one lexicographic sort with the line number as tie-break
    order = np.lexsort((np.arange(n), *split.cols[rel_col], *split.cols[ent_col]))
    a, b = split.ids[order, pa], split.ids[order, pb]
    new = np.ones(n, bool)
    new[1:] = (a[1:] != a[:-1]) | (b[1:] != b[:-1])
    starts = np.flatnonzero(new)
    line_ptr = np.concatenate([starts, [n]]).astype(np.int64)
    if drop_last_group and len(starts):
        line_ptr = line_ptr[:-1]                      # the final group is never flushed
    G = len(line_ptr) - 1
    used = order[: int(line_ptr[-1])] if G > 0 else order[:0]
    ptr, val = split.alt_ptr[ans_col], split.alt_val[ans_col]
    lens = ptr[used + 1] - ptr[used]
    alt_ptr = np.zeros(len(used) + 1, np.int64)
    np.cumsum(lens, out=alt_ptr[1:])
    src = np.repeat(ptr[used] - alt_ptr[:-1], lens) + np.arange(int(alt_ptr[-1]), dtype=np.int64)
    return Groups(np.stack([a[starts[:G]], b[starts[:G]]], axis=1) if G else np.zeros((0, 2), np.int64), slot,
                  line_ptr[: G + 1] if G else np.zeros(1, np.int64), alt_ptr, val[src])


def _group_value_rows(g: Groups) -> np.ndarray:
    """Group index of every flat mention id."""
    per_line = np.repeat(np.arange(len(g.line_ptr) - 1), np.diff(g.line_ptr))
    return np.repeat(per_line, np.diff(g.alt_ptr))


def merge_all_splits(groups: Sequence[Groups], exact_set_order: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """openkge/dataset.py:520-565 for one direction: union of the mention ids of every split per prefix; prefixes in
    numeric order. Returns (prefix [P, 2], ptr [P + 1], values)."""
    pa = np.concatenate([g.prefix[_group_value_rows(g), 0] for g in groups])
    pb = np.concatenate([g.prefix[_group_value_rows(g), 1] for g in groups])
    vals = np.concatenate([g.alt_val for g in groups])
    if len(vals) == 0:
        return np.zeros((0, 2), np.int64), np.zeros(1, np.int64), np.zeros(0, np.int64)
    order = np.lexsort((np.arange(len(vals)), pb, pa))             # stable: train, valid, test inside a prefix
    pa, pb, vals = pa[order], pb[order], vals[order]
    new = np.ones(len(vals), bool)
    new[1:] = (pa[1:] != pa[:-1]) | (pb[1:] != pb[:-1])
    starts = np.flatnonzero(new)
    prefix = np.stack([pa[starts], pb[starts]], axis=1)
    bounds = np.concatenate([starts, [len(vals)]])
    if exact_set_order:
        out_vals, ptr = [], [0]
        # the reference builds set(chain(*first line's lists)) and .update()s it line by line (one jsonl line per split)
        split_of = np.concatenate([np.full(len(g.alt_val), i) for i, g in enumerate(groups)])[order]
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            s: set = set()
            for sp in np.unique(split_of[lo:hi]):                  # np.unique sorts: train, valid, test
                chunk = vals[lo:hi][split_of[lo:hi] == sp].tolist()
                s = set(chunk) if not s else (s.update(chunk) or s)
            out_vals.extend(list(s))
            ptr.append(len(out_vals))
        return prefix, np.asarray(ptr, np.int64), np.asarray(out_vals, np.int64)
    # sorted unique values per prefix
    order2 = np.lexsort((vals, pb, pa))
    pa2, pb2, v2 = pa[order2], pb[order2], vals[order2]
    keep = np.ones(len(v2), bool)
    keep[1:] = (pa2[1:] != pa2[:-1]) | (pb2[1:] != pb2[:-1]) | (v2[1:] != v2[:-1])
    first = np.ones(len(v2), bool)
    first[1:] = (pa2[1:] != pa2[:-1]) | (pb2[1:] != pb2[:-1])
    counts = np.add.reduceat(keep.astype(np.int64), np.flatnonzero(first))
    ptr = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=ptr[1:])
    return prefix, ptr, v2[keep]


def pack_groups(g: Groups) -> Tuple[np.ndarray, np.ndarray]:
    """``pack_list_of_lists`` (utils/misc.py:56-75) of every group, concatenated: per group [k + 1 offsets, 0, values] with
    offsets relative to the group's segment. Returns (segment pointer [G + 1], packed int64 array)."""
    k = np.diff(g.line_ptr)                                        # answers per group
    G = len(k)
    line_group = np.repeat(np.arange(G), k)
    alt_len = np.diff(g.alt_ptr)
    v = np.add.reduceat(alt_len, g.line_ptr[:-1]) if G and len(alt_len) else np.zeros(G, np.int64)
    if G:
        v = np.where(k > 0, v, 0)
    seg_len = k + 2 + v
    seg_ptr = np.zeros(G + 1, np.int64)
    np.cumsum(seg_len, out=seg_ptr[1:])
    packed = np.zeros(int(seg_ptr[-1]), np.int64)
    # offsets: for answer j of group g: (k_g + 2) + #values of the group's earlier answers; plus the closing offset
    val_start_in_group = g.alt_ptr[:-1] - g.alt_ptr[g.line_ptr[:-1]][line_group]
    j_in_group = np.arange(len(line_group)) - g.line_ptr[:-1][line_group]
    packed[seg_ptr[:-1][line_group] + j_in_group] = k[line_group] + 2 + val_start_in_group
    packed[seg_ptr[:-1] + k] = k + 2 + v                          # closing offset; the 0 terminator follows (already 0)
    val_group = np.repeat(line_group, alt_len)
    val_pos = np.arange(len(g.alt_val)) - g.alt_ptr[g.line_ptr[:-1]][val_group]
    packed[seg_ptr[:-1][val_group] + k[val_group] + 2 + val_pos] = g.alt_val
    return seg_ptr, packed


def build_split_tensors(dataset_dir: str, input_file: str, train_file: str = "train.txt", valid_file: str = "valid.txt",
                        test_file: str = "test.txt", is_training_data: bool = True, max_size_prefix_label: int = -1,
                        drop_last_group: bool = True, exact_set_order: bool = False):
    """(seen_prefixes [P, 7] int32, seen_entities int32, all_splits_entities int32) of ``input_file`` exactly as
    ``create_data_tensors`` builds them (openkge/dataset.py:567-710)."""
    files = {f: read_split(os.path.join(dataset_dir, f)) for f in dict.fromkeys([train_file, valid_file, test_file, input_file])}
    rows, packed_parts, all_parts = [], [], []
    ent_off = all_off = 0
    for direction in _DIRECTIONS:
        per_split = {f: group_direction(s, direction, drop_last_group) for f, s in files.items()}
        a_prefix, a_ptr, a_val = merge_all_splits([per_split[train_file], per_split[valid_file], per_split[test_file]],
                                                  exact_set_order)
        g = per_split[input_file]
        if is_training_data and max_size_prefix_label > 1:
            g = chunk_groups(g, max_size_prefix_label)
        seg_ptr, packed = pack_groups(g)
        G = len(g.prefix)
        a0 = a1 = np.zeros(G, np.int64)
        if not is_training_data and G:
            # coordinates of the prefix in the all-splits lists (the dict lookup of :583-606)
            mult = np.int64(max(int(a_prefix[:, 1].max()), int(g.prefix[:, 1].max())) + 1)
            pos = np.searchsorted(a_prefix[:, 0] * mult + a_prefix[:, 1], g.prefix[:, 0] * mult + g.prefix[:, 1])
            a0, a1 = a_ptr[pos] + all_off, a_ptr[pos + 1] + all_off
        rows.append(np.stack([g.prefix[:, 0], g.prefix[:, 1], seg_ptr[:-1] + ent_off, seg_ptr[1:] + ent_off, a0, a1,
                              np.full(G, g.slot, np.int64)], axis=1) if G else np.zeros((0, 7), np.int64))
        packed_parts.append(packed)
        all_parts.append(a_val)
        ent_off += len(packed)
        all_off += len(a_val)
    return (np.concatenate(rows).astype(np.int32), np.concatenate(packed_parts).astype(np.int32),
            np.concatenate(all_parts).astype(np.int32))


def chunk_groups(g: Groups, max_size: int) -> Groups:
    """Rows of at most ``max_size`` answers for training (openkge/dataset.py:668-690); the line order is unchanged, so
    only the group boundaries move."""
    k = np.diff(g.line_ptr)
    n_chunks = np.where(k > max_size, -(-k // max_size), 1)
    rep = np.repeat(np.arange(len(k)), n_chunks)
    first = np.concatenate([[0], np.cumsum(n_chunks)[:-1]]) if len(k) else np.zeros(0, np.int64)
    within = np.arange(len(rep)) - first[rep]
    starts = g.line_ptr[:-1][rep] + within * max_size
    line_ptr = np.concatenate([starts, g.line_ptr[-1:]]).astype(np.int64)
    return Groups(g.prefix[rep], g.slot, line_ptr, g.alt_ptr, g.alt_val)


def load_meta(dataset_dir: str, entity_id_map: str = "entity_id_map.txt", relation_id_map: str = "relation_id_map.txt",
              entity_tokens: str = "entity_id_tokens_ids_map.txt", relation_tokens: str = "relation_id_tokens_ids_map.txt",
              max_length: Tuple[int, int] = (10, 10)) -> EntityRelationDatasetMeta:
    """``load_vocab`` (openkge/dataset.py:141-309): sizes are max id + 1; the token lists of the special ids 0 / 1 are
    ``[1]`` (:202-203); entities / relations without a token line do not occur (the reference asserts it)."""
    def max_id(path):
        m = 0
        with open(path, encoding="utf-8") as f:
            for i, line in enumerate(f):
                if i == 0 and line.startswith("#"):
                    continue
                m = max(m, int(line.split("\t")[1]))
        return m

    def token_lists(path, size):
        rows: List[Optional[List[int]]] = [None] * size
        tmax = 0
        with open(path, encoding="utf-8") as f:
            for i, line in enumerate(f):
                if i == 0 and line.startswith("#"):
                    continue
                key, toks = line.strip().split("\t")
                rows[int(key)] = [int(t) for t in toks.split()]
                tmax = max(tmax, max(rows[int(key)]))
        rows[0], rows[1] = [1], [1]
        assert all(r is not None for r in rows), "every id needs a token line (openkge/dataset.py:204-208)"
        return tuple(rows), tmax + 1

    meta = EntityRelationDatasetMeta(entities_size=max_id(os.path.join(dataset_dir, entity_id_map)) + 1,
                                     relations_size=max_id(os.path.join(dataset_dir, relation_id_map)) + 1,
                                     max_length=tuple(max_length))
    ep, rp = os.path.join(dataset_dir, entity_tokens), os.path.join(dataset_dir, relation_tokens)
    if os.path.exists(ep) and os.path.exists(rp):
        meta.entity_id_to_tokens_map, meta.entity_tokens_size = token_lists(ep, meta.entities_size)
        meta.relation_id_to_tokens_map, meta.relation_tokens_size = token_lists(rp, meta.relations_size)
    return meta


def load_prefix_index(dataset_dir: str, input_file: str, is_training_data: bool, meta: Optional[EntityRelationDatasetMeta] = None,
                      **kwargs) -> PrefixIndex:
    """Raw id files -> the decoded :class:`PrefixIndex` the B200 collate works on."""
    meta = meta if meta is not None else load_meta(dataset_dir)
    sp, se, ae = build_split_tensors(dataset_dir, input_file, is_training_data=is_training_data, **kwargs)
    return PrefixIndex(sp, se, ae, meta.entities_size, meta.min_entities_size, is_training_data)
