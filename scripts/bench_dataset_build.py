"""Dataset cache build: dataset_build.build_split_tensors vs the unmodified reference, on an FB15k-237-shaped graph written
in the reference's on-disk format (5-column id files + the six map files).

    python scripts/bench_dataset_build.py [n_train_triples]

Runs only where /root/reference exists (the build container); prints the timings and checks that the tensors are
bit-identical (exact_set_order=True) at this size too. CPU only."""
import os
import shutil
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_knowledge_graph_embeddings_b200 import dataset_build as B  # noqa: E402
from open_knowledge_graph_embeddings_b200 import synthetic as S  # noqa: E402


def write_dataset(root, spec, scale, seed=1):
    train, ev = S.make_triples(spec, seed, scale)
    half = len(ev) // 2
    splits = {"train.txt": train, "valid.txt": ev[:half], "test.txt": ev[half:]}
    for name, t in splits.items():
        with open(os.path.join(root, name), "w") as f:
            f.write("".join(f"{s}\t{r}\t{o}\t{s}\t{o}\n" for s, r, o in t.tolist()))
    meta = S.make_meta(spec, seed)
    for kind, n, rows, vocab in (("entity", spec.n_entities, meta.entity_token_rows, spec.entity_token_vocab),
                                 ("relation", spec.n_relations, meta.relation_token_rows, spec.relation_token_vocab)):
        with open(os.path.join(root, f"{kind}_id_map.txt"), "w") as f:
            f.write(f"# {kind}\tid\tcount\n" + "".join(f"x{i}\t{i}\t1\n" for i in range(2, n + 2)))
        with open(os.path.join(root, f"{kind}_token_id_map.txt"), "w") as f:
            f.write("# token\tid\tcount\n" + "".join(f"t{i}\t{i}\t1\n" for i in range(4, vocab + 4)))
        with open(os.path.join(root, f"{kind}_id_tokens_ids_map.txt"), "w") as f:
            f.write("# id\ttokens\n")
            for i in range(2, n + 2):
                toks = [int(t) for t in rows[i] if t]
                f.write(f"{i}\t{' '.join(map(str, toks))}\n")
            # make the largest token id appear so that *_tokens_size is the vocabulary size
        with open(os.path.join(root, f"{kind}_id_tokens_ids_map.txt"), "a") as f:
            pass
    return {k: len(v) for k, v in splits.items()}


def main():
    scale = float(sys.argv[1]) / 272115 if len(sys.argv) > 1 else 1.0
    spec = S.SPECS["fb15k237"]
    root = tempfile.mkdtemp(prefix="okge_ds_")
    try:
        sizes = write_dataset(root, spec, scale)
        print("dataset:", sizes)
        B._split_cache.clear()
        t0 = time.perf_counter()
        mine = {}
        for name, training in (("train.txt", True), ("valid.txt", False)):
            mine[name] = B.build_split_tensors(root, name, is_training_data=training, exact_set_order=True)
        t_mine_exact = time.perf_counter() - t0
        B._split_cache.clear()                                 # cold again: parsing the files is part of the job
        t0 = time.perf_counter()
        for name, training in (("train.txt", True), ("valid.txt", False)):
            B.build_split_tensors(root, name, is_training_data=training)
        t_mine = time.perf_counter() - t0
        print(f"dataset_build: {t_mine:.2f} s (sorted all-splits lists), {t_mine_exact:.2f} s (reference set order)")
        if not os.path.isdir("/root/reference"):
            print("reference not available here: no comparison")
            return
        sys.path.insert(0, "/root/reference")
        warnings.filterwarnings("ignore")
        import logging
        logging.disable(logging.CRITICAL)
        import tqdm
        tqdm.tqdm = lambda x, *a, **k: x                      # silence progress bars
        import openkge.dataset as RD
        RD.tqdm = lambda x, *a, **k: x
        common = dict(dataset_dir=root, loss="bce", replace_entities_by_tokens=True, replace_relations_by_tokens=True,
                      max_lengths_tuple=[10, 10], copy_data_to_dev_shm=False, device="cpu", batch_size=512)
        t0 = time.perf_counter()
        train = RD.OneToNMentionRelationDataset(input_file="train.txt", is_training_data=True, **common)
        valid = RD.OneToNMentionRelationDataset(input_file="valid.txt", is_training_data=False, **common)
        RD.OneToNMentionRelationDataset(input_file="test.txt", is_training_data=False, **common)
        valid.merge_all_splits_triples(root, "train.txt", "valid.txt", "test.txt")
        train.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
        valid.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
        t_ref = time.perf_counter() - t0
        print(f"reference (vocab load + prefix grouping + merge + tensors, first run, no caches): {t_ref:.2f} s")
        for name, ds in (("train.txt", train), ("valid.txt", valid)):
            sp, se, ae = mine[name]
            ok = (np.array_equal(sp, ds.seen_prefixes_tensor.numpy()) and np.array_equal(se, ds.seen_entities_tensor.numpy())
                  and np.array_equal(ae, ds.all_splits_entities_tensor.numpy()))
            print(f"{name}: {len(sp)} prefix rows, {len(se)} packed ints, {len(ae)} all-split ints, bit-identical: {ok}")
            assert ok
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
