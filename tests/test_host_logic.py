"""Host-side logic of the B200 path checked on CPU against the reference-generated fixtures:
wire-format helpers, the vectorised collate, the sparse batch types. No kernels are called."""
import os

import numpy as np
import pytest
import torch

from open_knowledge_graph_embeddings_b200 import dataset as D
from open_knowledge_graph_embeddings_b200.metrics import AccumulateMeter, MetricResult
from open_knowledge_graph_embeddings_b200.misc import pack_list_of_lists, unpack_list_of_lists


def test_pack_unpack_kats(kats):
    assert pack_list_of_lists([[5], [6, 7], [8]]) == kats["pack/a"].tolist()
    assert pack_list_of_lists([9, 10]) == kats["pack/b"].tolist()
    lol, flat = unpack_list_of_lists(kats["pack/a"])
    assert lol == [[5], [6, 7], [8]] and flat == kats["unpack/a_flat"].tolist()
    assert unpack_list_of_lists([]) == ([], [])


@pytest.mark.parametrize("split,training", [("train", True), ("valid", False)])
def test_prefix_index_collate_matches_reference(kats, split, training):
    sizes = kats["meta/sizes"]
    idx = D.PrefixIndex(kats[f"data/{split}/seen_prefixes"], kats[f"data/{split}/seen_entities"],
                        kats[f"data/{split}/all_splits_entities"], entity_vocab_size=int(sizes[0]),
                        entity_vocab_offset=2, is_training_data=training)
    assert len(idx) == len(kats[f"data/{split}/seen_prefixes"])
    slot_inputs, nl, nm, labels, label_ids, filt, shared = idx.collate(kats[f"collate/{split}/sampler"])
    po = torch.cat(slot_inputs[0], 1).numpy()
    sp = torch.cat(slot_inputs[1], 1).numpy()
    assert np.array_equal(po, kats[f"collate/{split}/po"]) and np.array_equal(sp, kats[f"collate/{split}/sp"])
    assert np.array_equal(labels.ptr.numpy(), kats[f"collate/{split}/pos_ptr"])
    assert np.array_equal(labels.idx.numpy(), kats[f"collate/{split}/pos_idx"])
    assert [nl, nm] == kats[f"collate/{split}/normalizers"].tolist()
    assert shared.shape == (int(sizes[0]) - 2, 1) and int(shared.materialize()[0]) == 2
    if training:
        assert label_ids is None and filt is None
    else:
        assert np.array_equal(filt.ptr.numpy(), kats[f"collate/{split}/filt_ptr"])
        assert np.array_equal(filt.idx.numpy(), kats[f"collate/{split}/filt_idx"])
        assert np.array_equal(label_ids.ans_row.numpy(), kats[f"collate/{split}/ans_row"])
        assert np.array_equal(label_ids.alt_ptr.numpy(), kats[f"collate/{split}/alt_ptr"])
        assert np.array_equal(label_ids.alt_idx.numpy(), kats[f"collate/{split}/alt_idx"])


def test_collate_edge_cases(kats):
    sizes = kats["meta/sizes"]
    idx = D.PrefixIndex(kats["data/valid/seen_prefixes"], kats["data/valid/seen_entities"],
                        kats["data/valid/all_splits_entities"], int(sizes[0]), 2, is_training_data=False)
    # only sp rows (stored first) -> po block is None, like the reference (openkge/dataset.py:887-889)
    slot_inputs, nl, nm, labels, label_ids, filt, _ = idx.collate([0, 1, 2])
    assert slot_inputs[0] is None and slot_inputs[1][0].shape == (3, 1)
    assert labels.shape == (3, int(sizes[0]) - 2) and nl == 3 * (int(sizes[0]) - 2)
    # empty index
    empty = D.PrefixIndex(np.zeros((0, 7), np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), 10, 2, False)
    assert len(empty) == 0


def test_csr_roundtrip_and_label_ids():
    dense = torch.zeros(4, 9)
    dense[0, [1, 7]] = 1
    dense[2, 3] = 1
    c = D.CSRMatrix.from_dense(dense)
    assert c.ptr.tolist() == [0, 2, 2, 3, 3] and c.idx.tolist() == [1, 7, 3]
    assert torch.equal(c.to_dense(), dense) and c.sum() == 3.0 and len(c) == 4 and c.size(1) == 9
    assert torch.equal(D.CSRMatrix.from_lists([[7, 1], [], [3], []], 9).to_dense(), dense)
    ra = D.RankedAnswers.from_label_ids([[torch.IntTensor([0]), torch.IntTensor([3, 5])], [torch.IntTensor([2])]])
    assert ra.ans_row.tolist() == [0, 0, 1] and ra.alt_ptr.tolist() == [0, 1, 3, 4] and ra.alt_idx.tolist() == [0, 3, 5, 2]


def test_meters_match_reference_semantics():
    m = AccumulateMeter()
    m.update(1.0, 1)
    m.update(0.0, 3)
    assert m.avg == 0.25 and m.count == 4
    a, b = MetricResult(), MetricResult()
    a["mrr"].update(0.5, 2)
    b["mrr"].update(1.0, 2)
    a = a + b
    assert a["mrr"].avg == 0.75 and a["mrr"].count == 4
    assert list(a.keys()) == ["loss", "h1", "h3", "h10", "h50", "mrr", "mr"]
    assert not a["loss"].greater_is_better and a["mrr"].greater_is_better


@pytest.mark.parametrize("split,training", [("train", True), ("valid", False)])
@pytest.mark.parametrize("variant,min_size", [("pad", 50), ("nopad", 0)])
def test_batch_shared_collate_matches_reference(kats, split, training, variant, min_size):
    """use_batch_shared_entities=True: candidate list (first-occurrence order + seeded negatives), local label /
    filter / answer columns — equal to the reference collate output (openkge/dataset.py:813-868, 899-919)."""
    sizes = kats["meta/sizes"]
    idx = D.PrefixIndex(kats[f"data/{split}/seen_prefixes"], kats[f"data/{split}/seen_entities"],
                        kats[f"data/{split}/all_splits_entities"], int(sizes[0]), 2, is_training_data=training)
    key = f"collate_shared/{split}/{variant}"
    np.random.seed(123)
    slot_inputs, nl, nm, labels, label_ids, filt, shared = D.collate_shared(idx, kats[f"{key}/sampler"], min_size)
    assert np.array_equal(shared.numpy().reshape(-1), kats[f"{key}/shared"])
    assert np.array_equal(torch.cat(slot_inputs[0], 1).numpy(), kats[f"{key}/po"])
    assert np.array_equal(torch.cat(slot_inputs[1], 1).numpy(), kats[f"{key}/sp"])
    assert np.array_equal(labels.ptr.numpy(), kats[f"{key}/pos_ptr"]) and np.array_equal(labels.idx.numpy(), kats[f"{key}/pos_idx"])
    assert [nl, nm] == kats[f"{key}/normalizers"].tolist()
    if not training:
        assert np.array_equal(filt.ptr.numpy(), kats[f"{key}/filt_ptr"]) and np.array_equal(filt.idx.numpy(), kats[f"{key}/filt_idx"])
        assert np.array_equal(label_ids.ans_row.numpy(), kats[f"{key}/ans_row"])
        assert np.array_equal(label_ids.alt_ptr.numpy(), kats[f"{key}/alt_ptr"])
        assert np.array_equal(label_ids.alt_idx.numpy(), kats[f"{key}/alt_idx"])


# ---------------------------------------------------------------------------------------------
# dataset cache build from the on-disk id files (openkge/dataset.py:141-309, 481-710)
# ---------------------------------------------------------------------------------------------

TINY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_dataset")


def test_dataset_build_is_bit_identical_to_the_reference_tensors(kats):
    """build_split_tensors on the reference-format files == the tensors dumped from the unmodified reference's
    create_data_tensors (string-ordered prefixes, dropped last group, CPython set order of the all-splits lists)."""
    from open_knowledge_graph_embeddings_b200 import dataset_build as B
    for split, fname, training in (("train", "train.txt", True), ("valid", "valid.txt", False)):
        sp, se, ae = B.build_split_tensors(TINY, fname, is_training_data=training, exact_set_order=True)
        assert sp.dtype == se.dtype == ae.dtype == np.int32
        assert np.array_equal(sp, kats[f"data/{split}/seen_prefixes"])
        assert np.array_equal(se, kats[f"data/{split}/seen_entities"])
        assert np.array_equal(ae, kats[f"data/{split}/all_splits_entities"])
    meta = B.load_meta(TINY)
    assert [meta.entities_size, meta.relations_size, meta.entity_tokens_size, meta.relation_tokens_size] == kats["meta/sizes"].tolist()
    assert meta.entity_id_to_tokens_map[0] == [1] and meta.entity_id_to_tokens_map[1] == [1]


def test_dataset_build_fast_path_and_index(kats):
    """Default (sorted) all-splits lists hold the same SETS as the reference's; the decoded PrefixIndex and its collate are
    identical to the ones built from the reference's own tensors; the last group really is the one the reference loses."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import dataset_build as B
    sizes = kats["meta/sizes"]
    sp, se, ae = B.build_split_tensors(TINY, "valid.txt", is_training_data=False)
    ref_sp, ref_ae = kats["data/valid/seen_prefixes"], kats["data/valid/all_splits_entities"]
    assert np.array_equal(sp[:, [0, 1, 2, 3, 6]], ref_sp[:, [0, 1, 2, 3, 6]])
    for row, ref_row in zip(sp, ref_sp):
        assert sorted(ae[row[4]:row[5]].tolist()) == sorted(ref_ae[ref_row[4]:ref_row[5]].tolist())
        assert ae[row[4]:row[5]].tolist() == sorted(ae[row[4]:row[5]].tolist())
    mine = B.load_prefix_index(TINY, "valid.txt", is_training_data=False)
    ref = D.PrefixIndex(ref_sp, kats["data/valid/seen_entities"], ref_ae, int(sizes[0]), 2, False)
    rows = np.arange(len(ref))
    a, b = mine.collate(rows), ref.collate(rows)
    for x, y in ((a[3], b[3]), (a[5], b[5])):                       # labels, filter
        assert torch.equal(x.ptr, y.ptr) and torch.equal(x.idx, y.idx)
    assert torch.equal(a[4].ans_row, b[4].ans_row) and torch.equal(a[4].alt_idx, b[4].alt_idx)
    # without the reference's lost-group bug every direction gains exactly one prefix
    full, _, _ = B.build_split_tensors(TINY, "valid.txt", is_training_data=False, drop_last_group=False)
    assert len(full) == len(sp) + 2


def test_dataset_build_chunks_long_training_groups():
    """max_size_prefix_label (openkge/dataset.py:668-690): groups with more answers become consecutive rows of at most that
    many; the concatenation of the chunks is the unchunked group; no phantom rows."""
    from open_knowledge_graph_embeddings_b200 import dataset_build as B
    from open_knowledge_graph_embeddings_b200.misc import unpack_list_of_lists
    plain_sp, plain_se, _ = B.build_split_tensors(TINY, "train.txt", is_training_data=True)
    sp, se, _ = B.build_split_tensors(TINY, "train.txt", is_training_data=True, max_size_prefix_label=2)
    assert len(sp) > len(plain_sp)
    groups, key_order = {}, []
    for a, b, s, e, _, _, slot in sp.tolist():
        lol, _ = unpack_list_of_lists(se[s:e])
        assert 1 <= len(lol) <= 2
        if (a, b, slot) not in groups:
            key_order.append((a, b, slot))
        groups.setdefault((a, b, slot), []).extend(lol)
    assert len(key_order) == len(plain_sp)
    for (a, b, s, e, _, _, slot), key in zip(plain_sp.tolist(), key_order):
        assert (a, b, slot) == key
        assert unpack_list_of_lists(plain_se[s:e])[0] == groups[key]


def test_ranked_answers_overflow_is_known_on_the_host():
    """RankedAnswers.overflow lists, per collated batch and without touching the device, the answers that are the 5th,
    6th, ... ranked answer of their prefix row (what the single-pass evaluation kernel cannot hold in its 4 slots)."""
    import torch
    from open_knowledge_graph_embeddings_b200.dataset import RankedAnswers
    rows = [0, 0, 1, 3, 3, 3, 3, 3, 3, 4]                       # row 3 has six answers: its 5th and 6th overflow
    ans = RankedAnswers(torch.tensor(rows, dtype=torch.int32), torch.arange(len(rows) + 1, dtype=torch.int32),
                        torch.arange(len(rows), dtype=torch.int32), slots=4)
    assert ans.overflow.dtype == torch.int64 and ans.overflow.tolist() == [7, 8]
    assert ans.extra_prefix.tolist() == [3] and ans.overflow_slot.tolist() == [0, 1]      # one extra query row for prefix 3
    assert RankedAnswers.from_label_ids([[torch.tensor([1])], [torch.tensor([2, 3])]]).overflow.numel() == 0
    pinned = ans.pin_memory() if torch.cuda.is_available() else ans
    assert pinned.overflow.tolist() == [7, 8] and pinned.extra_prefix.tolist() == [3]
    empty = RankedAnswers(torch.zeros(0, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), torch.zeros(0, dtype=torch.int32))
    assert empty.overflow.numel() == 0 and empty.extra_prefix.numel() == 0 and empty.nbytes == 4
    # a prefix with 11 answers needs two extra rows (4 + 4 + 3); another overflowing prefix comes after it
    rows = [2] * 11 + [5] * 6
    big = RankedAnswers(torch.tensor(rows, dtype=torch.int32), torch.arange(len(rows) + 1, dtype=torch.int32),
                        torch.arange(len(rows), dtype=torch.int32), slots=4)
    assert big.overflow.tolist() == [4, 5, 6, 7, 8, 9, 10, 15, 16]
    assert big.extra_prefix.tolist() == [2, 2, 5]
    assert big.overflow_slot.tolist() == [0, 1, 2, 3, 4, 5, 6, 8, 9]
    # slots chosen per batch: 512 prefix rows with one answer each and one row with three -> one slot per row and two
    # extra rows are cheaper ((13 + 4) * 5 tiles) than four slots for everybody ((13 + 16) * 4 tiles)
    rows = list(range(512)) + [511, 511]
    auto = RankedAnswers(torch.tensor(sorted(rows), dtype=torch.int32), torch.arange(len(rows) + 1, dtype=torch.int32),
                         torch.arange(len(rows), dtype=torch.int32), n_rows=512)
    assert auto.slots == 1 and auto.extra_prefix.tolist() == [511, 511] and auto.overflow.tolist() == [512, 513]
    assert auto.overflow_slot.tolist() == [0, 4]              # slot 0 of extra rows 0 and 1 (rows are 4 slots wide)


def test_split_parser_string_order_and_fallback(tmp_path):
    """The vectorised split parser: (1) its numeric keys sort exactly like Python's comparison of the decimal STRINGS (the
    order the reference groups prefixes in); (2) it returns the same Split as the line-by-line parser; (3) files it does
    not accept (a sign, a blank line) take the text path."""
    import numpy as np
    from open_knowledge_graph_embeddings_b200 import dataset_build as B
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.integers(0, 10 ** rng.integers(1, 8), 400), [0, 1, 10, 100, 12, 120, 1200, 9, 99, 999]]).astype(np.int64)
    order = np.lexsort((np.arange(v.size), *B._string_order_keys(v)))
    expect = sorted(range(v.size), key=lambda i: (str(int(v[i])), i))
    assert order.tolist() == expect
    lines = []
    for _ in range(300):
        s_, r_, o_ = (int(x) for x in rng.integers(2, 500, 3))
        alts = lambda x: " ".join(str(int(a)) for a in [x] + list(rng.integers(2, 500, int(rng.integers(0, 4)))))   # noqa: E731
        lines.append(f"{s_}\t{r_}\t{o_}\t{alts(s_)}\t{alts(o_)}\n")
    path = tmp_path / "train.txt"
    path.write_text("".join(lines))
    fast, slow = B._parse_split(str(path)), B._parse_split_text(str(path))
    assert len(fast.cols[0]) == 2 and len(slow.cols[0]) == 1          # numeric keys vs strings
    assert np.array_equal(fast.ids, slow.ids)
    for c in (3, 4):
        assert np.array_equal(fast.alt_ptr[c], slow.alt_ptr[c]) and np.array_equal(fast.alt_val[c], slow.alt_val[c])
    for d in B._DIRECTIONS:
        gf, gs = B.group_direction(fast, d), B.group_direction(slow, d)
        assert np.array_equal(gf.prefix, gs.prefix) and np.array_equal(gf.line_ptr, gs.line_ptr)
        assert np.array_equal(gf.alt_ptr, gs.alt_ptr) and np.array_equal(gf.alt_val, gs.alt_val)
    odd = tmp_path / "odd.txt"
    odd.write_text("".join(lines[:5]) + "\n" + "+7\t3\t4\t7\t4\n")
    parsed = B._parse_split(str(odd))
    assert len(parsed.cols[0]) == 1 and parsed.ids[-1].tolist() == [7, 3, 4] and len(parsed.ids) == 6


def test_collate_many_equals_collate_and_loader_uses_it(kats):
    """dataset.collate_many (k training batches in one vectorised pass, batches are views of the big tensors) returns
    exactly what k collate calls return, and get_loader yields the same batches whether or not it chunks."""
    import numpy as np
    import torch
    from open_knowledge_graph_embeddings_b200 import dataset as D
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)

    def same(one, other):
        for s1, s2 in zip(one[0], other[0]):
            assert (s1 is None) == (s2 is None)
            if s1 is not None:
                assert all(torch.equal(x, y) and x.shape == y.shape and x.dtype == y.dtype and y.is_contiguous() for x, y in zip(s1, s2))
        assert one[1] == other[1] and one[2] == other[2] and one[3].shape == other[3].shape
        assert torch.equal(one[3].ptr, other[3].ptr) and torch.equal(one[3].idx, other[3].idx)
        assert other[3].ptr.is_contiguous() and other[3].idx.is_contiguous()

    rng = np.random.default_rng(1)
    rows = rng.integers(0, len(tr_idx), (5, 16))
    rows[0] = np.flatnonzero(tr_idx.slot == 0)[:16]                  # a batch without sp rows
    rows[1] = np.flatnonzero(tr_idx.slot == 2)[:16]                  # ... and one without po rows
    for r, m in zip(rows, D.collate_many(tr_idx, rows)):
        same(tr_idx.collate(r), m)
    ds = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=16, device="cpu", is_training_data=True)
    chunked = list(ds.get_loader(shuffle=True, drop_last=False, pin_memory=False, seed=4))
    order = D.LazyPermutation(len(tr_idx), 4)[0:len(tr_idx)]         # the row order of epoch 0 of get_loader(seed=4)
    assert len(chunked) == (len(tr_idx) + 15) // 16
    for i, b in enumerate(chunked):
        same(tr_idx.collate(order[16 * i:16 * (i + 1)]), b)
    prefetched = list(ds.get_loader(shuffle=True, drop_last=False, pin_memory=False, seed=4, prefetch=2))
    assert len(prefetched) == len(chunked)
    for a, b in zip(chunked, prefetched):                            # background thread: same batches, same order
        same(a, b)
    ev_idx = D.PrefixIndex(kats["data/valid/seen_prefixes"], kats["data/valid/seen_entities"],
                           kats["data/valid/all_splits_entities"], int(sizes[0]), 2, False)
    with pytest.raises(ValueError):
        D.collate_many(ev_idx, rows % len(ev_idx))


def test_collate_shared_fast_sampling_keeps_the_batch_semantics(kats):
    """exact_sampling=False (no full permutation of all entity ids per batch): the candidate list still starts with the
    batch's own answers in first-occurrence order, is topped up to min_size_batch_labels with distinct other entities,
    and labels are the same (row, entity) pairs as with the reference-exact sampling; the lookup table is handed back clean."""
    import numpy as np
    from open_knowledge_graph_embeddings_b200 import dataset as D
    sizes = kats["meta/sizes"]
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    rows = np.random.default_rng(2).integers(0, len(tr_idx), 24)
    np.random.seed(5)
    exact = D.collate_shared(tr_idx, rows, 50)
    np.random.seed(5)
    fast = D.collate_shared(tr_idx, rows, 50, exact_sampling=False)
    e_ids, f_ids = exact[6].reshape(-1).numpy(), fast[6].reshape(-1).numpy()
    assert len(f_ids) == len(e_ids) == 50 and len(set(f_ids.tolist())) == 50
    n_keys = int(exact[3].idx.max()) + 1                               # answers occupy the first positions
    assert np.array_equal(f_ids[:n_keys], e_ids[:n_keys])
    assert f_ids.min() >= 2 and f_ids.max() < int(sizes[0])

    def pairs(batch):
        ptr, idx, ids = batch[3].ptr.numpy(), batch[3].idx.numpy(), batch[6].reshape(-1).numpy()
        return sorted((int(r), int(ids[c])) for r in range(len(ptr) - 1) for c in idx[ptr[r]:ptr[r + 1]])
    assert pairs(fast) == pairs(exact)
    assert (tr_idx.__dict__["_shared_lut"] == -1).all()


def test_lazy_permutation_is_a_permutation_and_reshuffles():
    """The shuffled loader's row order (dataset.LazyPermutation): every slice-wise evaluation of an epoch visits each row
    exactly once, different seeds give different orders, slices agree with the whole."""
    from open_knowledge_graph_embeddings_b200.dataset import LazyPermutation
    for n in (1, 2, 3, 17, 1000, 4097, 65536, 100003):
        p = LazyPermutation(n, seed=n)
        whole = p[0:n]
        assert whole.dtype == np.int64 and np.array_equal(np.sort(whole), np.arange(n))
        parts = np.concatenate([p[i:i + 333] for i in range(0, n, 333)])
        assert np.array_equal(parts, whole)
    a, b = LazyPermutation(50000, 1)[0:50000], LazyPermutation(50000, 2)[0:50000]
    assert (a == b).mean() < 0.01 and abs(np.corrcoef(a, np.arange(50000))[0, 1]) < 0.02
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.02


# ---------------------------------------------------------------------------------------------
# the reference's REAL FB15k-237 fixture (tests/golden/real_fixture.py; goldens from the unmodified reference)
# ---------------------------------------------------------------------------------------------

def test_dataset_build_on_the_real_fb15k237_fixture_matches_the_reference(tmp_path):
    """build_split_tensors on the reference's own FB15k-237 id files (17,535 / 10,000 / 10,466 triples, 14,541 entities,
    237 relations, real token maps) produces byte-identical tensors to the reference's create_data_tensors for all three
    splits (SHA-256 of seen_prefixes / seen_entities / all_splits_entities) and the same vocabulary sizes."""
    import hashlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import real_fixture
    from open_knowledge_graph_embeddings_b200 import dataset_build as B
    gold = np.load(os.path.join(os.path.dirname(real_fixture.__file__), "real_fb15k237.npz"))
    root = real_fixture.stage(str(tmp_path / "fb"))
    for split, fname, training in (("train", "train.txt", True), ("valid", "valid.txt", False), ("test", "test.txt", False)):
        tensors = B.build_split_tensors(root, fname, is_training_data=training, exact_set_order=True)
        for name, t in zip(("seen_prefixes", "seen_entities", "all_splits_entities"), tensors):
            assert list(t.shape) == gold[f"data/{split}/{name}/shape"].tolist(), (split, name)
            digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(t).tobytes()).digest(), np.uint8)
            assert np.array_equal(digest, gold[f"data/{split}/{name}/sha256"]), (split, name)
    meta = B.load_meta(root)
    assert [meta.entities_size, meta.relations_size, meta.entity_tokens_size, meta.relation_tokens_size] == \
        gold["meta/sizes"].tolist() == [14543, 239, 17324, 452]



def test_collate_many_packed_payload_matches_the_tuple():
    """dataset.PackedBatch: the 7-tuple a loader yields also carries its integer payload as ONE int32 array
    [ent | rel | ptr, n_po | idx] (dataset.packed_layout) -- what a graphed step copies to the device in a single H2D copy.
    It must say exactly what the tuple says, and the tuple must still be the reference's wire format."""
    from tests.conftest import load_golden
    from open_knowledge_graph_embeddings_b200 import dataset as D
    kats = load_golden("kats")
    sizes = kats["meta/sizes"]
    idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                        kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    rng = np.random.default_rng(0)
    for B in (32, 7):
        rows = rng.integers(0, len(idx), (5, B))
        o_ent, o_rel, o_ptr, o_idx = D.packed_layout(B)
        assert o_ent == 0 and o_rel % 4 == 0 and o_ptr % 4 == 0 and o_idx % 4 == 0 and o_idx - o_ptr >= B + 2
        for b, r in zip(D.collate_many(idx, rows), rows):
            assert len(b) == 7 and len(tuple(b)) == 7 and b[1] == B * idx.n_cols and b[2] == float(b[3].idx.numel())
            ref = idx.collate(r)
            po, sp = b[0]
            ent = torch.cat([t for t in ((po[1] if po else None), (sp[0] if sp else None)) if t is not None]).reshape(-1)
            rel = torch.cat([t for t in ((po[0] if po else None), (sp[1] if sp else None)) if t is not None]).reshape(-1)
            pk = b.packed
            assert pk.dtype == torch.int32 and pk.is_contiguous()
            assert torch.equal(pk[:B], ent.int()) and torch.equal(pk[o_rel:o_rel + B], rel.int())
            assert torch.equal(pk[o_ptr:o_ptr + B + 1], b[3].ptr) and int(pk[o_ptr + B + 1]) == b.n_po
            assert torch.equal(pk[o_idx:o_idx + b[3].idx.numel()], b[3].idx) and pk.numel() == o_idx + b[3].idx.numel()
            assert torch.equal(ref[3].idx, b[3].idx) and torch.equal(ref[3].ptr, b[3].ptr)
            assert b.n_po == (0 if po is None else po[0].numel())


def test_union_slot_numbering_of_the_touched_row_exchange():
    """functional.touch_flags / union_slots (plain tensor code, runs anywhere): the token rows any rank touches get the slots
    base, base + 1, ... in ascending row order, the PAD token never counts, rows nobody touches map to -1, and a count
    above the capacity spills into the dump row on the write side and reads as 'no gradient' (the graphed step always
    picks a capacity that holds the count; this is the out-of-bounds guard behind it)."""
    from open_knowledge_graph_embeddings_b200 import functional as Fn
    id_rows = torch.tensor([[2, 5, 7, 3, 0, 0], [2, 9, 3, 0, 0, 0], [2, 11, 12, 13, 3, 0], [1, 0, 0, 0, 0, 0]], dtype=torch.int32)
    V = 16
    flags_a, flags_b = torch.zeros(V, dtype=torch.int32), torch.zeros(V, dtype=torch.int32)
    Fn.touch_flags(flags_a, id_rows, torch.tensor([0, 1], dtype=torch.int32))          # rank 0 looks up rows 0, 1
    Fn.touch_flags(flags_b, id_rows, torch.tensor([2], dtype=torch.int32))             # rank 1 looks up row 2
    assert flags_a.tolist() == [0, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0] and flags_b[0] == 0 and flags_b[11] == 1
    union = torch.maximum(flags_a, flags_b)                                           # what the all-reduce(max) leaves
    wm, rm = torch.full((V,), -7, dtype=torch.int32), torch.full((V,), -7, dtype=torch.int32)
    nxt = Fn.union_slots(union, torch.zeros((), dtype=torch.int64), 100, wm, rm)
    touched = [2, 3, 5, 7, 9, 11, 12, 13]
    assert int(nxt) == len(touched) and torch.equal(wm, rm)
    assert [int(wm[t]) for t in touched] == list(range(len(touched)))
    assert all(int(wm[t]) == -1 for t in range(V) if t not in touched)
    # a second table continues the numbering; capacity 10 < 8 + 4: the overflow writes to the dump row (slot 10), reads -1
    flags2 = torch.tensor([0, 1, 0, 1, 1, 1], dtype=torch.int32)
    wm2, rm2 = torch.zeros(6, dtype=torch.int32), torch.zeros(6, dtype=torch.int32)
    nxt2 = Fn.union_slots(flags2, nxt, 10, wm2, rm2)
    assert int(nxt2) == 12 and wm2.tolist() == [-1, 8, -1, 9, 10, 10] and rm2.tolist() == [-1, 8, -1, 9, -1, -1]


def test_host_collate_entry_points_edge_cases():
    """okge_host_collate_plan / _fill (plain C, no GPU): batches that are all po or all sp rows, rows without labels, a
    one-row batch, and the error path (a row index outside the prefix table is refused, nothing is written)."""
    from tests.conftest import load_golden
    from open_knowledge_graph_embeddings_b200 import _capi
    from open_knowledge_graph_embeddings_b200 import dataset as D
    kats = load_golden("kats")
    sizes = kats["meta/sizes"]
    idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                        kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    po_rows, sp_rows = np.flatnonzero(idx.slot == 0), np.flatnonzero(idx.slot == 2)
    assert len(po_rows) >= 6 and len(sp_rows) >= 6
    for rows in (po_rows[:6].reshape(2, 3), sp_rows[:6].reshape(2, 3), np.array([[po_rows[0]], [sp_rows[0]]])):
        for b, r in zip(D.collate_many(idx, rows), rows):
            ref = idx.collate(r)
            assert b.n_po == int((idx.slot[r] == 0).sum()) and b.rows == len(r)
            (po, sp), (rpo, rsp) = b[0], ref[0]
            for got, want in ((po, rpo), (sp, rsp)):
                assert (got is None) == (want is None)
                if got is not None:
                    assert all(torch.equal(g.int(), w.int()) for g, w in zip(got, want))
            assert torch.equal(b[3].ptr, ref[3].ptr) and torch.equal(b[3].idx, ref[3].idx) and b[2] == ref[2] and b[1] == ref[1]
    bad = np.array([[0, len(idx)]], dtype=np.int64)                     # the second index is one past the table
    with pytest.raises(_capi.OkgeNativeError, match="out of range"):
        D.collate_many(idx, bad)
