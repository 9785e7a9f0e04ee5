"""Seeded synthetic knowledge graphs in the shapes of the benchmark configs.

There is no network for datasets, so throughput runs use synthetic triples with the statistics of the
real ones: Zipf-distributed entity / relation frequencies, prefix groups built exactly like the
reference builds them (group train triples by (s, r) and by (r, o), openkge/dataset.py:481-518; filter
sets are the union over all splits, :520-565), optional alternative-mention lists and token rows.
Everything is vectorised numpy so that a 10 M-triple graph is indexed in seconds.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .dataset import EntityRelationDatasetMeta, PrefixIndex

OFFSET = 2  # ids 0 / 1 are PAD / UNK (openkge/index_mapper.py:14)


@dataclass
class GraphSpec:
    name: str
    n_entities: int
    n_relations: int
    n_train: int
    n_eval: int
    zipf_a: float = 1.0
    max_alternatives: int = 1          # > 1: OLPBench-style alternative answer mentions
    entity_token_vocab: int = 0        # > 0: token model
    relation_token_vocab: int = 0
    max_size_prefix_label: int = -1    # > 1: training answer lists split into chunks of that many (dataset.py:668-690)


SPECS = {
    # fb15k237-complex-kge.yaml / -unigrampool.yaml shapes (14,541 entities, 237 relations, 272,115 train)
    "fb15k237": GraphSpec("fb15k237", 14541, 237, 272115, 17535, entity_token_vocab=17320, relation_token_vocab=448),
    # LookupDistmult 1-vs-all, 1 M entities
    "c3_1m": GraphSpec("c3_1m", 1_000_000, 1000, 10_000_000, 10_000),
    # OLPBench-shaped: 2.5 M mentions, ~1 M relations, 30 M triples, token vocab 200 k / 50 k
    "olpbench": GraphSpec("olpbench", 2_500_000, 1_000_000, 30_000_000, 10_000, max_alternatives=10,
                          entity_token_vocab=200_000, relation_token_vocab=50_000, max_size_prefix_label=64),
}


def _zipf_ids(rng: np.random.Generator, n_ids: int, size: int, a: float) -> np.ndarray:
    """ids in [OFFSET, OFFSET + n_ids) with P(rank k) ~ 1 / k^a (inverse-CDF sampling, vectorised)."""
    ranks = np.arange(1, n_ids + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-a))
    cdf /= cdf[-1]
    u = rng.random(size)
    idx = np.searchsorted(cdf, u, side="left")
    perm = rng.permutation(n_ids)       # popularity is not tied to the id order
    return (perm[np.minimum(idx, n_ids - 1)] + OFFSET).astype(np.int64)


def _sorted_unique(x: np.ndarray) -> np.ndarray:
    """np.unique for 1-D int64 via sort + adjacent compare (numpy 2.3's unique is ~70x slower here)."""
    x = np.sort(x)
    if len(x) == 0:
        return x
    keep = np.ones(len(x), bool)
    keep[1:] = x[1:] != x[:-1]
    return x[keep]


def make_triples(spec: GraphSpec, seed: int = 1, scale: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """(train [n, 3], eval [m, 3]) int64 (s, r, o) triples, duplicates removed (via one composite int64 key
    per triple: 1-D sort instead of a row-wise unique)."""
    rng = np.random.default_rng(seed)
    n = int(spec.n_train * scale) + spec.n_eval
    s = _zipf_ids(rng, spec.n_entities, n, spec.zipf_a)
    r = _zipf_ids(rng, spec.n_relations, n, spec.zipf_a)
    o = _zipf_ids(rng, spec.n_entities, n, spec.zipf_a)
    E, R = np.int64(spec.n_entities + OFFSET), np.int64(spec.n_relations + OFFSET)
    key = _sorted_unique((s * R + r) * E + o)
    rng.shuffle(key)
    t = np.stack([key // (R * E), (key // E) % R, key % E], axis=1)
    return t[spec.n_eval:], t[:spec.n_eval]


def _group(keys_a: np.ndarray, keys_b: np.ndarray, vals: np.ndarray, base_b: int, base_v: int):
    """Group by (a, b): returns unique prefixes [P, 2], CSR (ptr, sorted unique vals). One 1-D sort of the
    composite key (a * base_b + b) * base_v + v."""
    key = _sorted_unique((keys_a * np.int64(base_b) + keys_b) * np.int64(base_v) + vals)
    pref = key // np.int64(base_v)
    v = key % np.int64(base_v)
    new_group = np.ones(len(key), bool)
    new_group[1:] = pref[1:] != pref[:-1]
    starts = np.flatnonzero(new_group)
    ptr = np.concatenate([starts, [len(v)]]).astype(np.int64)
    p = pref[starts]
    return np.stack([p // np.int64(base_b), p % np.int64(base_b)], axis=1), ptr, v


def _lookup_groups(prefix_sorted_keys: np.ndarray, ptr: np.ndarray, vals: np.ndarray, query_keys: np.ndarray):
    """CSR rows of `query_keys` (int64 composite keys) in a grouped structure; missing -> empty."""
    pos = np.searchsorted(prefix_sorted_keys, query_keys)
    pos = np.minimum(pos, len(prefix_sorted_keys) - 1)
    hit = prefix_sorted_keys[pos] == query_keys
    lens = np.where(hit, ptr[pos + 1] - ptr[pos], 0)
    out_ptr = np.zeros(len(query_keys) + 1, np.int64)
    np.cumsum(lens, out=out_ptr[1:])
    starts = np.repeat(ptr[pos] - out_ptr[:-1], lens)
    return out_ptr, vals[starts + np.arange(int(out_ptr[-1]), dtype=np.int64)]


def _chunk_rows(prefix: np.ndarray, slot: np.ndarray, ptr: np.ndarray, max_len: int):
    """Training rows with more than `max_len` answers become ceil(len / max_len) rows of the same prefix, each with
    the next `max_len` answers (openkge/dataset.py:668-690; wikiopenlink-thorough-complex-lstm.yaml:156 sets 64). The
    answers stay where they are, only the row boundaries are refined."""
    lens = np.diff(ptr)
    n_chunks = np.maximum((lens + max_len - 1) // max_len, 1)
    row = np.repeat(np.arange(len(lens), dtype=np.int64), n_chunks)
    first = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(n_chunks, out=first[1:])
    k = np.arange(int(first[-1]), dtype=np.int64) - first[row]
    new_ptr = np.concatenate([ptr[row] + k * max_len, ptr[-1:]]).astype(np.int64)
    return prefix[row], slot[row], new_ptr


def build_indexes(spec: GraphSpec, seed: int = 1, scale: float = 1.0):
    """(train PrefixIndex, eval PrefixIndex, meta). Train rows: one per distinct (s, r) [slot 2] and (r, o)
    [slot 0] prefix with all its training answers as labels. Eval rows: the prefixes of the eval triples,
    labels = eval answers, filter = answers known in train + eval."""
    train, ev = make_triples(spec, seed, scale)
    n_cols = spec.n_entities
    mult = np.int64(spec.n_entities + spec.n_relations + OFFSET + 1)

    def both_directions(t):
        E, R = spec.n_entities + OFFSET, spec.n_relations + OFFSET
        sp_pref, sp_ptr, sp_val = _group(t[:, 0], t[:, 1], t[:, 2], R, E)   # (s, r) -> o   slot 2
        po_pref, po_ptr, po_val = _group(t[:, 1], t[:, 2], t[:, 0], E, E)   # (r, o) -> s   slot 0
        return (sp_pref, sp_ptr, sp_val), (po_pref, po_ptr, po_val)

    def assemble(sp, po, filt=None, alternatives=None):
        prefix = np.concatenate([sp[0], po[0]]).astype(np.int32)
        slot = np.concatenate([np.full(len(sp[0]), 2, np.int32), np.full(len(po[0]), 0, np.int32)])
        lab_ptr = np.concatenate([sp[1], po[1][1:] + sp[1][-1]])
        lab_idx = (np.concatenate([sp[2], po[2]]) - OFFSET).astype(np.int32)
        return prefix, slot, lab_ptr, lab_idx

    tr_sp, tr_po = both_directions(train)
    prefix, slot, lab_ptr, lab_idx = assemble(tr_sp, tr_po)
    if spec.max_size_prefix_label > 1:
        prefix, slot, lab_ptr = _chunk_rows(prefix, slot, lab_ptr, spec.max_size_prefix_label)
    tr = PrefixIndex.from_csr(prefix, slot, lab_ptr, lab_idx, n_cols=n_cols, offset=OFFSET, is_training_data=True)

    ev_sp, ev_po = both_directions(ev)
    all_sp, all_po = both_directions(np.concatenate([train, ev]))
    prefix, slot, lab_ptr, lab_idx = assemble(ev_sp, ev_po)
    f_sp = _lookup_groups(all_sp[0][:, 0] * mult + all_sp[0][:, 1], all_sp[1], all_sp[2],
                          ev_sp[0][:, 0] * mult + ev_sp[0][:, 1])
    f_po = _lookup_groups(all_po[0][:, 0] * mult + all_po[0][:, 1], all_po[1], all_po[2],
                          ev_po[0][:, 0] * mult + ev_po[0][:, 1])
    filt_ptr = np.concatenate([f_sp[0], f_po[0][1:] + f_sp[0][-1]])
    filt_idx = (np.concatenate([f_sp[1], f_po[1]]) - OFFSET).astype(np.int32)
    # ranked answers: one per (prefix, eval answer); alternative mentions: the answer itself plus, for
    # OLPBench-shaped graphs, up to max_alternatives - 1 extra mention ids
    rng = np.random.default_rng(seed + 17)
    n_ans = len(lab_idx)
    if spec.max_alternatives > 1:
        extra = rng.integers(0, spec.max_alternatives, n_ans)
        alt_ptr = np.zeros(n_ans + 1, np.int64)
        np.cumsum(1 + extra, out=alt_ptr[1:])
        alt_idx = rng.integers(0, n_cols, int(alt_ptr[-1])).astype(np.int32)
        alt_idx[alt_ptr[:-1]] = lab_idx
    else:
        alt_ptr = np.arange(n_ans + 1, dtype=np.int64)
        alt_idx = lab_idx.copy()
    evi = PrefixIndex.from_csr(prefix, slot, lab_ptr, lab_idx, n_cols=n_cols, offset=OFFSET, is_training_data=False,
                               ans_ptr=lab_ptr, alt_ptr=alt_ptr, alt_idx=alt_idx, filt_ptr=filt_ptr, filt_idx=filt_idx)
    meta = make_meta(spec, seed)
    return tr, evi, meta


def token_rows(rng: np.random.Generator, n_rows: int, vocab: int, mean_len: float, max_len: int = 10) -> np.ndarray:
    """[n_rows + 2, max_len] int64 token-id rows in the layout of TokenBasedRelationEmbedder
    (openkge/model.py:576-595): BOS=2 ... EOS=3, last max_len tokens, left aligned, PAD=0 filled; rows 0 / 1
    (PAD / UNK entity) are [1, 0, ...] (openkge/dataset.py:202-203)."""
    body = np.clip(rng.geometric(1.0 / max(mean_len - 2.0, 1.0), n_rows), 1, max_len - 2)
    rows = np.zeros((n_rows + OFFSET, max_len), np.int64)
    rows[:OFFSET, 0] = 1
    toks = rng.integers(4, vocab + 4, (n_rows, max_len - 2))
    col = np.arange(max_len)[None, :]
    length = (body + 2)[:, None]
    real = rows[OFFSET:]
    real[:, 0] = 2
    inner = (col >= 1) & (col < length - 1)
    real[:, 1:max_len - 1] = np.where(inner[:, 1:max_len - 1], toks, 0)
    real[np.arange(n_rows), body + 1] = 3
    return rows


def make_meta(spec: GraphSpec, seed: int = 1) -> EntityRelationDatasetMeta:
    meta = EntityRelationDatasetMeta(entities_size=spec.n_entities + OFFSET, relations_size=spec.n_relations + OFFSET)
    if spec.entity_token_vocab:
        rng = np.random.default_rng(seed + 5)
        meta.entity_tokens_size = spec.entity_token_vocab + 4
        meta.relation_tokens_size = spec.relation_token_vocab + 4
        meta.entity_token_rows = token_rows(rng, spec.n_entities, spec.entity_token_vocab, 4.8)
        meta.relation_token_rows = token_rows(rng, spec.n_relations, spec.relation_token_vocab, 8.0)
    return meta
