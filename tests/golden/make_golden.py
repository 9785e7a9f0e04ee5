"""Generate golden vectors by running the UNMODIFIED reference (read-only at /root/reference).

Only runs in the build container (the reference does not travel to the GPU box); the outputs are
the small ``*.npz`` fixtures committed next to this script. Re-run with

    python tests/golden/make_golden.py

What it does: writes a tiny synthetic dataset in the reference's on-disk format (5-column id
triples + the six map files), builds the reference datasets / collate / models /
AddLossModule / OptimRegime / compute_metrics from ``openkge`` and ``utils`` exactly as
``scripts/train.py`` wires them (that script itself crashes in ResultsLog on pandas >= 2), and dumps
inputs, scores, loss, gradients, post-step weights and rank counts for every model on the hot path.

Work-arounds applied OUTSIDE the reference tree: ``model.entity_projection = None``
for UnigramPooling (attribute is never defined, openkge/model.py:789), relation_slot_size ==
entity_slot_size, dropout = 0 (RNG streams cannot match), seeding done here.
"""
import os
import random
import shutil
import sys
import tempfile
import warnings

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from openkge.dataset import OneToNMentionRelationDataset  # noqa: E402
from openkge.model import Models  # noqa: E402
from openkge.trainer import AddLossModule  # noqa: E402
from utils.misc import pack_list_of_lists, unpack_list_of_lists  # noqa: E402
from utils.optim import OptimRegime  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

N_ENT, N_REL = 61, 7            # real entities / relations (ids 2 .. N+1)
ENT_TOK_V, REL_TOK_V = 40, 12   # real token vocab sizes (ids 4 .. V+3)


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def write_dataset(root, rng):
    """Synthetic KG with alternative mentions in columns 4/5 (OLPBench style)."""
    os.makedirs(root, exist_ok=True)
    ent_ids = np.arange(2, N_ENT + 2)
    rel_ids = np.arange(2, N_REL + 2)

    def triples(n):
        rows = set()
        while len(rows) < n:
            s = int(rng.choice(ent_ids[:25]))  # few subjects -> multi-answer prefixes
            r = int(rng.choice(rel_ids))
            o = int(rng.choice(ent_ids))
            rows.add((s, r, o))
        return sorted(rows)

    all_t = triples(260)
    rng.shuffle(all_t)
    splits = {"train.txt": all_t[:180], "valid.txt": all_t[180:220], "test.txt": all_t[220:]}
    for name, rows in splits.items():
        with open(os.path.join(root, name), "w") as f:
            for (s, r, o) in rows:
                s_alt = sorted({s} | ({int(rng.choice(ent_ids))} if rng.random() < 0.3 else set()))
                o_alt = sorted({o} | ({int(rng.choice(ent_ids))} if rng.random() < 0.3 else set()))
                f.write(f"{s}\t{r}\t{o}\t{' '.join(map(str, s_alt))}\t{' '.join(map(str, o_alt))}\n")

    with open(os.path.join(root, "entity_id_map.txt"), "w") as f:
        f.write("# entity\tid\tcount\n")
        for i in ent_ids:
            f.write(f"e{i}\t{i}\t1\n")
    with open(os.path.join(root, "relation_id_map.txt"), "w") as f:
        f.write("# relation\tid\tcount\n")
        for i in rel_ids:
            f.write(f"r{i}\t{i}\t1\n")
    with open(os.path.join(root, "entity_token_id_map.txt"), "w") as f:
        f.write("# token\tid\tcount\n")
        for i in range(4, ENT_TOK_V + 4):
            f.write(f"t{i}\t{i}\t1\n")
    with open(os.path.join(root, "relation_token_id_map.txt"), "w") as f:
        f.write("# token\tid\tcount\n")
        for i in range(4, REL_TOK_V + 4):
            f.write(f"t{i}\t{i}\t1\n")
    # token rows incl. BOS=2 / EOS=3; some longer than 10 so the "last 10" truncation is exercised
    with open(os.path.join(root, "entity_id_tokens_ids_map.txt"), "w") as f:
        f.write("# id\ttokens\n")
        for i in ent_ids:
            n = int(rng.integers(1, 13))
            toks = [2] + [int(t) for t in rng.integers(4, ENT_TOK_V + 4, size=n)] + [3]
            f.write(f"{i}\t{' '.join(map(str, toks))}\n")
        # force the maximum token id to appear so entity_tokens_size is deterministic
    with open(os.path.join(root, "relation_id_tokens_ids_map.txt"), "w") as f:
        f.write("# id\ttokens\n")
        for i in rel_ids:
            n = int(rng.integers(1, 14))
            toks = [2] + [int(t) for t in rng.integers(4, REL_TOK_V + 4, size=n)] + [3]
            f.write(f"{i}\t{' '.join(map(str, toks))}\n")


def build_datasets(root, batch_size):
    common = dict(dataset_dir=root, loss="bce", replace_entities_by_tokens=True, replace_relations_by_tokens=True,
                  max_lengths_tuple=[10, 10], copy_data_to_dev_shm=False, device="cpu", batch_size=batch_size)
    train = OneToNMentionRelationDataset(input_file="train.txt", is_training_data=True, **common)
    valid = OneToNMentionRelationDataset(input_file="valid.txt", is_training_data=False, **common)
    test = OneToNMentionRelationDataset(input_file="test.txt", is_training_data=False, **common)
    valid.merge_all_splits_triples(root, "train.txt", "valid.txt", "test.txt")
    train.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
    valid.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
    return train, valid, test


def mixed_sampler(ds, n):
    """Indices that mix sp rows (stored first) and po rows (stored last) of seen_prefixes_tensor, so one
    batch holds both prefix kinds like a shuffled loader would produce (openkge/dataset.py:794-811)."""
    P = len(ds)
    idx = []
    for i in range(n):
        idx.append(i // 2 if i % 2 == 0 else P - 1 - i // 2)
    return idx


def sparse_rows(dense):
    """[B, N] 0/1 -> CSR (ptr, idx)."""
    ptr, idx = [0], []
    for row in dense:
        nz = np.nonzero(row)[0]
        idx.extend(nz.tolist())
        ptr.append(len(idx))
    return np.asarray(ptr, np.int32), np.asarray(idx, np.int32)


def flatten_label_ids(label_ids):
    """list[B] of list of IntTensor -> (ans_row [Q], alt_ptr [Q+1], alt_idx)."""
    ans_row, alt_ptr, alt_idx = [], [0], []
    for b, labels in enumerate(label_ids):
        for alt in labels:
            ans_row.append(b)
            alt_idx.extend(alt.tolist())
            alt_ptr.append(len(alt_idx))
    return np.asarray(ans_row, np.int32), np.asarray(alt_ptr, np.int32), np.asarray(alt_idx, np.int32)


def reference_rank_counts(filter_mask, label_ids, predictions):
    """The inner quantities of compute_metrics (openkge/dataset.py:430-445), recomputed with the
    reference's own tensor expressions so that they can be stored next to the meter averages."""
    true, greater, equal = [], [], []
    for prefix_filter, prefix_labels, prefix_prediction in zip(filter_mask, label_ids, predictions):
        rep = prefix_prediction.unsqueeze(0).repeat(len(prefix_labels), 1)
        frep = prefix_filter.unsqueeze(0).repeat(len(prefix_labels), 1)
        tl = [prefix_prediction[l.long()].max(0)[0] for l in prefix_labels]
        t = torch.Tensor(tl)
        rep.masked_fill_(frep, -1e8)
        greater.extend((t.view(len(prefix_labels), -1) < rep).long().sum(1).tolist())
        equal.extend((t.view(len(prefix_labels), -1) == rep).long().sum(1).tolist())
        true.extend(t.tolist())
    return np.asarray(true, np.float32), np.asarray(greater, np.int64), np.asarray(equal, np.int64)


def run_model_case(name, model_name, model_config, loss_name, smoothing, train, valid, root, optimizer):
    seed_all(7)
    meta = train.get_dataset_meta_dict()
    model = getattr(Models, model_name)(**model_config, train_data=meta)
    if model_name.startswith("UnigramPooling"):
        model.entity_projection = None
    loss = torch.nn.BCEWithLogitsLoss(reduction="sum") if loss_name == "bce" else torch.nn.KLDivLoss(reduction="sum")
    mwl = AddLossModule(model, loss, smoothing)
    args = {"optimization_config": dict(optimizer), "lr_scheduler_config": None}
    optimizers = OptimRegime.setup_optimizer_regime(args=args, model=model)

    out = {}
    for k, v in model.state_dict().items():
        out["init/" + k] = v.detach().numpy().copy()

    # ---- one training step on the first (unshuffled) batch ----
    loader = train.get_loader(sampler=mixed_sampler(train, train.batch_size), num_workers=0, drop_last=True)
    batch = next(iter(loader))
    out["train/sampler"] = np.asarray(mixed_sampler(train, train.batch_size), np.int64)
    inputs, nl, nm, labels, _, _, shared = train.input_and_labels_to_device(batch, training=True, device="cpu")
    out["train/po_rel"], out["train/po_obj"] = (t.numpy().reshape(-1).copy() for t in inputs[0])
    out["train/sp_subj"], out["train/sp_rel"] = (t.numpy().reshape(-1).copy() for t in inputs[1])
    out["train/pos_ptr"], out["train/pos_idx"] = sparse_rows(labels.numpy())
    out["train/normalizer_loss"] = np.asarray(nl, np.int64)
    out["train/normalizer_metric"] = np.asarray(nm, np.float64)
    out["train/shared_ids"] = shared.numpy().reshape(-1).copy()

    model.train()
    for o in optimizers:
        o.update(1, 0)
        o.zero_grad()
    loss_v, hook, scores = mwl(inputs=inputs, labels=labels.clone(), batch_shared_entities=shared,
                               use_batch_shared_entities=False, epoch=1,
                               input_style_triple_or_prefix="right_and_left_prefix")
    assert hook is None
    (loss_v.sum() / nl).backward()
    out["train/scores"] = scores.detach().numpy().copy()
    out["train/loss_sum"] = np.asarray(loss_v.detach().item(), np.float64)
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.detach().numpy().copy()
    for o in optimizers:
        o.step()
    for k, v in model.state_dict().items():
        out["step1/" + k] = v.detach().numpy().copy()
    for o in optimizers:
        for p_name, p in model.named_parameters():
            st = o.optimizer.state.get(p, {})
            for sk, sv in st.items():
                if torch.is_tensor(sv) and sv.dim() > 0:
                    out[f"optstate/{p_name}/{sk}"] = sv.detach().numpy().copy()
        grp = o.optimizer.param_groups[0]
        out["opt/eps"] = np.asarray(grp["eps"], np.float64)
        out["opt/lr"] = np.asarray(grp["lr"], np.float64)
        out["opt/weight_decay"] = np.asarray(grp["weight_decay"], np.float64)

    # ---- evaluation of the first validation batch with the updated model ----
    model.eval()
    vloader = valid.get_loader(sampler=mixed_sampler(valid, valid.batch_size), num_workers=0, drop_last=False)
    vbatch = next(iter(vloader))
    out["eval/sampler"] = np.asarray(mixed_sampler(valid, valid.batch_size), np.int64)
    inputs, nl, nm, labels, label_ids, filt, shared = valid.input_and_labels_to_device(vbatch, training=False, device="cpu")
    with torch.no_grad():
        loss_v, _, scores = mwl(inputs=inputs, labels=labels.clone(), batch_shared_entities=shared,
                                use_batch_shared_entities=False, epoch=1,
                                input_style_triple_or_prefix="right_and_left_prefix")
    metrics = OneToNMentionRelationDataset.compute_metrics(filt, label_ids, scores)
    out["eval/po_rel"], out["eval/po_obj"] = (t.numpy().reshape(-1).copy() for t in inputs[0])
    out["eval/sp_subj"], out["eval/sp_rel"] = (t.numpy().reshape(-1).copy() for t in inputs[1])
    out["eval/pos_ptr"], out["eval/pos_idx"] = sparse_rows(labels.numpy())
    out["eval/filt_ptr"], out["eval/filt_idx"] = sparse_rows(filt.numpy())
    out["eval/ans_row"], out["eval/alt_ptr"], out["eval/alt_idx"] = flatten_label_ids(label_ids)
    out["eval/scores"] = scores.numpy().copy()
    out["eval/loss_sum"] = np.asarray(loss_v.item(), np.float64)
    out["eval/normalizer_loss"] = np.asarray(nl, np.int64)
    t, g, e = reference_rank_counts(filt, label_ids, scores)
    out["eval/true_score"], out["eval/greater"], out["eval/equal"] = t, g, e
    for k, m in metrics.items():
        out[f"eval/metric/{k}"] = np.asarray([m.avg, m.count], np.float64)

    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: loss {out['train/loss_sum']:.6f} eval mrr {metrics['mrr'].avg:.6f} "
          f"({int(metrics['mrr'].count)} ranked answers)")


def run_trajectory_case(name, model_name, model_config, train, valid, optimizer, steps=30):
    """A short TRAINING RUN of the unmodified reference: `steps` optimizer steps over seeded shuffled batches (the loss of
    every step), then the filtered evaluation of the whole validation split (MRR / Hits). Pins the trajectory, not just one
    step: a port whose per-step error compounds would drift off this curve."""
    seed_all(11)
    meta = train.get_dataset_meta_dict()
    model = getattr(Models, model_name)(**model_config, train_data=meta)
    if model_name.startswith("UnigramPooling"):
        model.entity_projection = None
    mwl = AddLossModule(model, torch.nn.BCEWithLogitsLoss(reduction="sum"), 0.0)
    optimizers = OptimRegime.setup_optimizer_regime(args={"optimization_config": dict(optimizer), "lr_scheduler_config": None},
                                                    model=model)
    out = {}
    for k, v in model.state_dict().items():
        out["init/" + k] = v.detach().numpy().copy()
    rng = np.random.default_rng(5)
    bs = train.batch_size
    rows, losses = [], []
    model.train()
    step = 0
    while step < steps:
        order = rng.permutation(len(train))
        for i in range(0, len(order) - bs + 1, bs):
            if step >= steps:
                break
            sel = order[i:i + bs].tolist()
            batch = next(iter(train.get_loader(sampler=sel, num_workers=0, drop_last=True)))
            inputs, nl, nm, labels, _, _, shared = train.input_and_labels_to_device(batch, training=True, device="cpu")
            step += 1
            for o in optimizers:
                o.update(1, step)
                o.zero_grad()
            loss_v, hook, _ = mwl(inputs=inputs, labels=labels.clone(), batch_shared_entities=shared,
                                  use_batch_shared_entities=False, epoch=1, input_style_triple_or_prefix="right_and_left_prefix")
            (loss_v.sum() / nl).backward()
            for o in optimizers:
                o.step()
            rows.append(sel)
            losses.append(loss_v.detach().item() / nl)
    out["traj/rows"] = np.asarray(rows, np.int64)
    out["traj/loss"] = np.asarray(losses, np.float64)
    for k, v in model.state_dict().items():
        out["final/" + k] = v.detach().numpy().copy()
    # filtered evaluation of the whole validation split, in file order
    model.eval()
    from utils.metrics import MetricResult
    total = MetricResult()
    with torch.no_grad():
        for vbatch in valid.get_loader(shuffle=False, num_workers=0, drop_last=False):
            inputs, nl, nm, labels, label_ids, filt, shared = valid.input_and_labels_to_device(vbatch, training=False, device="cpu")
            _, _, scores = mwl(inputs=inputs, labels=labels.clone(), batch_shared_entities=shared,
                               use_batch_shared_entities=False, epoch=1, input_style_triple_or_prefix="right_and_left_prefix")
            total = total + OneToNMentionRelationDataset.compute_metrics(filt, label_ids, scores)
    for k, m in total.items():
        if k != "loss":
            out[f"eval/metric/{k}"] = np.asarray([m.avg, m.count], np.float64)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: loss {losses[0]:.6f} -> {losses[-1]:.6f} after {steps} steps, eval mrr {total['mrr'].avg:.6f} "
          f"({int(total['mrr'].count)} ranked answers)")


def run_kats(train, valid, root):
    out = {}
    # pack / unpack (utils/misc.py:56-89)
    out["pack/a"] = np.asarray(pack_list_of_lists([[5], [6, 7], [8]]), np.int64)
    out["pack/b"] = np.asarray(pack_list_of_lists([9, 10]), np.int64)
    lol, flat = unpack_list_of_lists(torch.IntTensor(pack_list_of_lists([[5], [6, 7], [8]])))
    out["unpack/a_flat"] = np.asarray(flat, np.int64)
    out["unpack/a_lens"] = np.asarray([len(x) for x in lol], np.int64)
    # compute_metrics KAT of SURVEY.md §4
    scores = torch.tensor([[.5, .9, .5, .1, .5, .7, .5, .3], [1.] * 8])
    filt = torch.zeros(2, 8).bool()
    filt[0, [0, 5]] = True
    filt[1, 2] = True
    label_ids = [[torch.IntTensor([0]), torch.IntTensor([3, 5])], [torch.IntTensor([2])]]
    m = OneToNMentionRelationDataset.compute_metrics(filt, label_ids, scores)
    t, g, e = reference_rank_counts(filt, label_ids, scores)
    out["kat/scores"] = scores.numpy()
    out["kat/filt_ptr"], out["kat/filt_idx"] = sparse_rows(filt.numpy())
    out["kat/ans_row"], out["kat/alt_ptr"], out["kat/alt_idx"] = flatten_label_ids(label_ids)
    out["kat/true"], out["kat/greater"], out["kat/equal"] = t, g, e
    for k, mm in m.items():
        out[f"kat/metric/{k}"] = np.asarray([mm.avg, mm.count], np.float64)
    # dataset tensors (wire format of the collate input, openkge/dataset.py:567-710)
    for nm_, ds in (("train", train), ("valid", valid)):
        out[f"data/{nm_}/seen_prefixes"] = ds.seen_prefixes_tensor.numpy().copy()
        out[f"data/{nm_}/seen_entities"] = ds.seen_entities_tensor.numpy().copy()
        out[f"data/{nm_}/all_splits_entities"] = ds.all_splits_entities_tensor.numpy().copy()
    meta = train.get_dataset_meta_dict()
    out["meta/sizes"] = np.asarray([meta.entities_size, meta.relations_size, meta.entity_tokens_size,
                                    meta.relation_tokens_size], np.int64)
    # collate outputs for the first batch of each split, dense form
    for nm_, ds, tr in (("train", train, True), ("valid", valid, False)):
        b = next(iter(ds.get_loader(sampler=mixed_sampler(ds, ds.batch_size), num_workers=0, drop_last=False)))
        inputs, nl, nmet, labels, label_ids, filt, shared = b
        out[f"collate/{nm_}/sampler"] = np.asarray(mixed_sampler(ds, ds.batch_size), np.int64)
        out[f"collate/{nm_}/po"] = torch.cat(inputs[0], 1).numpy()
        out[f"collate/{nm_}/sp"] = torch.cat(inputs[1], 1).numpy()
        out[f"collate/{nm_}/pos_ptr"], out[f"collate/{nm_}/pos_idx"] = sparse_rows(labels.numpy())
        out[f"collate/{nm_}/normalizers"] = np.asarray([nl, nmet], np.float64)
        if not tr:
            out[f"collate/{nm_}/filt_ptr"], out[f"collate/{nm_}/filt_idx"] = sparse_rows(filt.numpy())
            a, p, i = flatten_label_ids(label_ids)
            out[f"collate/{nm_}/ans_row"], out[f"collate/{nm_}/alt_ptr"], out[f"collate/{nm_}/alt_idx"] = a, p, i
    # batch-shared-entities collate (openkge/dataset.py:813-868) with negative sampling, seeded numpy generator
    common = dict(dataset_dir=root, loss="bce", replace_entities_by_tokens=True, replace_relations_by_tokens=True,
                  max_lengths_tuple=[10, 10], copy_data_to_dev_shm=False, device="cpu", batch_size=24,
                  use_batch_shared_entities=True, min_size_batch_labels=50)
    tr_s = OneToNMentionRelationDataset(input_file="train.txt", is_training_data=True, **common)
    va_s = OneToNMentionRelationDataset(input_file="valid.txt", is_training_data=False, **common)
    tr_s.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
    va_s.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
    for nm_, ds, tr in (("train", tr_s, True), ("valid", va_s, False)):
        for variant, min_size in (("pad", 50), ("nopad", 0)):
            ds.min_size_batch_labels = min_size
            np.random.seed(123)
            b = next(iter(ds.get_loader(sampler=mixed_sampler(ds, ds.batch_size), num_workers=0, drop_last=False)))
            inputs, nl, nmet, labels, label_ids, filt, shared = b
            key = f"collate_shared/{nm_}/{variant}"
            out[f"{key}/sampler"] = np.asarray(mixed_sampler(ds, ds.batch_size), np.int64)
            out[f"{key}/po"] = torch.cat(inputs[0], 1).numpy()
            out[f"{key}/sp"] = torch.cat(inputs[1], 1).numpy()
            out[f"{key}/shared"] = shared.numpy().reshape(-1)
            out[f"{key}/pos_ptr"], out[f"{key}/pos_idx"] = sparse_rows(labels.numpy())
            out[f"{key}/normalizers"] = np.asarray([nl, nmet], np.float64)
            if not tr:
                out[f"{key}/filt_ptr"], out[f"{key}/filt_idx"] = sparse_rows(filt.numpy())
                a, p_, i = flatten_label_ids(label_ids)
                out[f"{key}/ans_row"], out[f"{key}/alt_ptr"], out[f"{key}/alt_idx"] = a, p_, i
    np.savez_compressed(os.path.join(OUT, "kats.npz"), **out)
    print("wrote kats.npz")


def run_real_fixture_case():
    """The reference's REAL FB15k-237 fixture (data/fb15k237/mapped_to_ids; staged by real_fixture.stage because its
    train.txt is absent): digests + shapes of the tensors create_data_tensors builds for every split, and the filtered
    evaluation of the whole validation split (10,000 triples -> ~20 k ranked answers against 14,541 entities) with the
    unmodified reference's fp32 CPU scorer for a lookup and a token model whose weights both sides draw from the same seeded
    numpy generator: per-answer (greater, equal) counts and the meter averages."""
    import hashlib
    sys.path.insert(0, OUT)
    import real_fixture
    from utils.metrics import MetricResult
    src = os.path.join(REF, "data", "fb15k237", "mapped_to_ids")
    real_fixture.make_archive(src)
    root = tempfile.mkdtemp(prefix="okge_real_")
    try:
        real_fixture.stage(root, src)
        train, valid, test = build_datasets(root, batch_size=256)
        test.create_data_tensors(root, "train.txt", "valid.txt", "test.txt")
        out = {}
        for nm_, ds in (("train", train), ("valid", valid), ("test", test)):
            for tname in ("seen_prefixes", "seen_entities", "all_splits_entities"):
                t = getattr(ds, tname + "_tensor").numpy()
                out[f"data/{nm_}/{tname}/sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(t).tobytes()).digest(), np.uint8)
                out[f"data/{nm_}/{tname}/shape"] = np.asarray(t.shape, np.int64)
        meta = train.get_dataset_meta_dict()
        out["meta/sizes"] = np.asarray([meta.entities_size, meta.relations_size, meta.entity_tokens_size,
                                        meta.relation_tokens_size], np.int64)
        cases = (("lookup_complex", "LookupComplexRelationModel", dict(entity_slot_size=64, init_std=0.1, sparse=False)),
                 ("unigram_complex", "UnigramPoolingComplexRelationModel",
                  dict(entity_slot_size=64, relation_slot_size=64, init_std=0.1, sparse=False, pool="sum", dropout=0.0)))
        for cname, model_name, cfg in cases:
            seed_all(3)
            model = getattr(Models, model_name)(**cfg, train_data=meta)
            if model_name.startswith("UnigramPooling"):
                model.entity_projection = None
            floats = [(k, tuple(v.shape)) for k, v in model.state_dict().items() if v.dtype.is_floating_point]
            weights = real_fixture.seeded_weights(floats)
            model.load_state_dict({k: torch.from_numpy(w) for k, w in weights.items()}, strict=False)
            out[f"{cname}/keys"] = np.asarray(sorted(k for k, _ in floats))
            out[f"{cname}/shapes"] = np.asarray([list(sh) + [0] * (2 - len(sh)) for _, sh in sorted(floats)], np.int64)
            mwl = AddLossModule(model, torch.nn.BCEWithLogitsLoss(reduction="sum"), 0.0)
            model.eval()
            if hasattr(model, "entity_token_ids"):             # token models: cached encode of every entity
                with torch.no_grad():
                    model.precompute_embeddings_from_tokens()
            total, greater, equal, true = MetricResult(), [], [], []
            with torch.no_grad():
                for vbatch in valid.get_loader(shuffle=False, num_workers=0, drop_last=False):
                    inputs, nl, nm, labels, label_ids, filt, shared = valid.input_and_labels_to_device(vbatch, training=False,
                                                                                                       device="cpu")
                    _, _, scores = mwl(inputs=inputs, labels=labels.clone(), batch_shared_entities=shared,
                                       use_batch_shared_entities=False, epoch=1,
                                       input_style_triple_or_prefix="right_and_left_prefix")
                    total = total + OneToNMentionRelationDataset.compute_metrics(filt, label_ids, scores)
                    t, g, e = reference_rank_counts(filt, label_ids, scores)
                    true.append(t), greater.append(g), equal.append(e)
            out[f"{cname}/greater"] = np.concatenate(greater).astype(np.int32)
            out[f"{cname}/equal"] = np.concatenate(equal).astype(np.int32)
            out[f"{cname}/true"] = np.concatenate(true).astype(np.float32)
            for k, m in total.items():
                if k != "loss":
                    out[f"{cname}/metric/{k}"] = np.asarray([m.avg, m.count], np.float64)
            print(f"real fixture {cname}: {int(total['mrr'].count)} ranked answers, mrr {total['mrr'].avg:.6f} "
                  f"h10 {total['h10'].avg:.4f} mr {total['mr'].avg:.1f}")
        np.savez_compressed(os.path.join(OUT, "real_fb15k237.npz"), **out)
        print("wrote real_fb15k237.npz and fb15k237_ids.tar.gz")
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main():
    if sys.argv[1:] == ["real_fixture"]:             # only the real-fixture goldens (reads /root/reference/data)
        return run_real_fixture_case()
    rng = np.random.default_rng(20240607)
    root = tempfile.mkdtemp(prefix="okge_golden_")
    try:
        write_dataset(root, rng)
        # ship the tiny dataset itself so that the new data path can be tested on the same files
        ds_out = os.path.join(OUT, "tiny_dataset")
        shutil.rmtree(ds_out, ignore_errors=True)
        os.makedirs(ds_out)
        for f in ("train.txt", "valid.txt", "test.txt", "entity_id_map.txt", "relation_id_map.txt",
                  "entity_token_id_map.txt", "relation_token_id_map.txt", "entity_id_tokens_ids_map.txt",
                  "relation_id_tokens_ids_map.txt"):
            shutil.copy(os.path.join(root, f), ds_out)
        train, valid, test = build_datasets(root, batch_size=48)
        only = set(sys.argv[1:])             # `make_golden.py lstm_complex_bce ...` regenerates just the named cases
        if only:
            global run_model_case
            full_run = run_model_case
            run_model_case = lambda name, *a, **k: full_run(name, *a, **k) if name in only else None  # noqa: E731
        if not only or "kats" in only:
            run_kats(train, valid, root)
        adagrad = {"optimizer": "Adagrad", "epoch": 0, "lr": 0.3, "weight_decay": 1e-10}
        adam = {"optimizer": "Adam", "epoch": 0, "lr": 0.01}
        lookup = dict(entity_slot_size=32, init_std=0.1, sparse=False)
        run_model_case("lookup_distmult_bce", "LookupDistmultRelationModel", lookup, "bce", 0.0, train, valid, root, adagrad)
        run_model_case("lookup_complex_bce", "LookupComplexRelationModel", lookup, "bce", 0.0, train, valid, root, adagrad)
        run_model_case("lookup_complex_bce_smooth", "LookupComplexRelationModel", lookup, "bce", 0.1, train, valid, root, adagrad)
        run_model_case("lookup_complex_kl_adam", "LookupComplexRelationModel", lookup, "kl", 0.0, train, valid, root, adam)
        for pool in ("sum", "mean", "max"):
            uni = dict(entity_slot_size=32, relation_slot_size=32, init_std=0.1, sparse=False, pool=pool, dropout=0.0)
            run_model_case(f"unigram_complex_{pool}_bce", "UnigramPoolingComplexRelationModel", uni, "bce", 0.0, train,
                           valid, root, adagrad)
        uni_bn = dict(entity_slot_size=32, relation_slot_size=32, init_std=0.1, sparse=False, pool="sum", dropout=0.0,
                      normalize="batchnorm")
        run_model_case("unigram_complex_sum_bn_bce", "UnigramPoolingComplexRelationModel", uni_bn, "bce", 0.0, train,
                       valid, root, adagrad)
        # LSTM encoders (openkge/model.py:912-998), the OLPBench headline family; dropout 0 (RNG streams cannot match)
        lstm = dict(entity_slot_size=32, relation_slot_size=32, init_std=0.1, sparse=False, dropout=0.0)
        run_model_case("lstm_complex_bce", "LSTMComplexRelationModel", lstm, "bce", 0.0, train, valid, root, adagrad)
        run_model_case("lstm_distmult_bn_bce", "LSTMDistmultRelationModel", dict(lstm, normalize="batchnorm"), "bce", 0.0,
                       train, valid, root, adagrad)
        # 30-step training trajectories + final filtered evaluation of the whole validation split
        if not only or "traj_lookup_complex" in only:
            run_trajectory_case("traj_lookup_complex", "LookupComplexRelationModel", lookup, train, valid, adagrad)
        if not only or "traj_unigram_bn" in only:
            run_trajectory_case("traj_unigram_bn", "UnigramPoolingComplexRelationModel", uni_bn, train, valid, adagrad)
        if not only:
            run_real_fixture_case()
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
