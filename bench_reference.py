"""Reference arm of bench.py: the UNMODIFIED reference (samuelbroscheit/open_knowledge_graph_embeddings) timed on the
host cores of the GPU box.

The reference is pure Python without a setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference`
fails ("Neither 'setup.py' nor 'pyproject.toml' found"). `install()` — called by `__graft_entry__.build()` in the build
container, where /root/reference exists — copies its two Python packages (`openkge/`, `utils/`) verbatim into
`baseline/_ref/` (git-ignored, shipped to the GPU box with the working tree); this module puts that directory on
sys.path and drives the reference's OWN classes: `Models.*`, `AddLossModule.forward`, `loss.backward()`,
`OptimRegime.step()` (training) and `OneToNMentionRelationDataset.compute_metrics` (evaluation) — the hot path of
openkge/trainer.py:181-272 — on batches collated in the reference's wire format (dense fp32 [B, N] labels, dense bool
filter mask, list-of-lists label ids). Nothing of this repo's model / kernel code is on that path; only the synthetic
graph generator and the (bit-equal to the reference, see tests/test_host_logic.py) collate are shared, because the
reference's own dataset class needs its on-disk text format and minutes of Python loops to build a 1 M-entity index.
"""
import os
import shutil
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
REF_SRC = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "openkge", "trainer.py")) and os.path.isfile(os.path.join(REF_DIR, "utils", "optim.py"))


def install() -> str:
    """Copy the reference's Python packages into baseline/_ref (no-op without /root/reference). Returns a status line."""
    if not os.path.isdir(REF_SRC):
        return "no /root/reference here: baseline/_ref left as it is (" + ("present" if available() else "absent") + ")"
    for pkg in ("openkge", "utils"):
        dst = os.path.join(REF_DIR, pkg)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(REF_SRC, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return f"copied {REF_SRC}/{{openkge,utils}} to {REF_DIR}"


def _import_reference():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    warnings.filterwarnings("ignore")
    from openkge.dataset import EntityRelationDatasetMeta, OneToNMentionRelationDataset  # noqa: E402
    from openkge.model import Models  # noqa: E402
    from openkge.trainer import AddLossModule  # noqa: E402
    from utils.optim import OptimRegime  # noqa: E402
    return EntityRelationDatasetMeta, OneToNMentionRelationDataset, Models, AddLossModule, OptimRegime


def _build_model(wl, meta, seed):
    """The reference's model class for the workload, random init, on CPU."""
    RefMeta, _, Models, AddLossModule, OptimRegime = _import_reference()
    token = "Unigram" in wl["model"] or "LSTM" in wl["model"]
    rmeta = RefMeta(entity_id_count_map={}, relation_id_count_map={}, entity_token_id_count_map={}, relation_token_id_count_map={},
                    # token rows are assigned below as buffers: the reference builds them with a Python loop per entity
                    entity_id_to_tokens_map=[[1]] if token else {}, relation_id_to_tokens_map=[[1]] if token else {},
                    entities_size=meta.entities_size, relations_size=meta.relations_size, min_entities_size=2,
                    min_relations_size=2, entity_tokens_size=getattr(meta, "entity_tokens_size", 0),
                    relation_tokens_size=getattr(meta, "relation_tokens_size", 0), max_length=(10, 10))
    torch.manual_seed(seed)
    cfg = dict(wl["model_config"])
    model = getattr(Models, wl["model"])(entity_slot_size=wl["dim"], train_data=rmeta, **cfg)
    if token:
        model.entity_token_ids = torch.from_numpy(np.asarray(meta.entity_token_rows)).long()
        model.relation_token_ids = torch.from_numpy(np.asarray(meta.relation_token_rows)).long()
        if "Unigram" in wl["model"]:
            model.entity_projection = None        # never defined by the reference (openkge/model.py:789), SURVEY 8c (3)
    loss = torch.nn.BCEWithLogitsLoss(reduction="sum")
    mwl = AddLossModule(model, loss, 0.0)
    args = {"optimization_config": {"optimizer": "Adagrad", "epoch": 0, "lr": wl["lr"], "weight_decay": wl["weight_decay"]},
            "lr_scheduler_config": None}
    opts = OptimRegime.setup_optimizer_regime(args=args, model=model)
    return model, mwl, opts


def _dense_wire_batch(batch, n_cols, training):
    """A sparse batch of this repo's collate in the reference's wire format (openkge/dataset.py:937-940)."""
    slot_inputs, nl, nm, labels, ans, filt, shared = batch
    y = labels.to_dense(torch.float32)
    label_ids = filter_mask = None
    if not training:
        filter_mask = filt.to_dense(torch.float32).bool()
        ar, ap, ai = ans.ans_row.numpy(), ans.alt_ptr.numpy(), ans.alt_idx
        label_ids = [[ai[ap[j]:ap[j + 1]] for j in np.flatnonzero(ar == b)] for b in range(len(labels))]
    if shared is None or not isinstance(shared, torch.Tensor):
        shared = torch.arange(2, n_cols + 2, dtype=torch.int32).unsqueeze(1)       # openkge/dataset.py:872
    return slot_inputs, nl, nm, y, label_ids, filter_mask, shared


def train_run(workloads, workload, steps, warmup, budget_s, seed=1):
    """Reference training steps on synthetic batches of the workload. Returns (triples/s, ms/step, cores, sample, rows)."""
    from open_knowledge_graph_embeddings_b200 import dataset as DS
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    wl = workloads[workload]
    spec = S.SPECS[wl["spec"]]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)                                              # openkge/trainer.py:136
    tr_idx, _, meta = S.build_indexes(spec, seed=seed, scale=wl.get("scale", 1.0))
    np.random.seed(seed)
    model, mwl, opts = _build_model(wl, meta, seed)
    model.train()
    B = wl["batch"]
    shared_mode = wl.get("shared", False)
    n_cols = []
    state = {"i": 0, "steps": 0}

    def one_step(b):
        rng = np.random.default_rng(seed + 100 + state["i"])
        state["i"] += 1
        rows = rng.integers(0, len(tr_idx), b)
        if shared_mode:
            batch = DS.collate_shared(tr_idx, rows, wl.get("min_size_batch_labels", -1))
        else:
            batch = tr_idx.collate(rows)
        inputs, nl, nm, y, _, _, shared = _dense_wire_batch(batch, spec.n_entities, True)
        n_cols.append(y.shape[1])
        t0 = time.perf_counter()
        state["steps"] += 1
        for o in opts:
            o.update(1, state["steps"])
            o.zero_grad()
        loss, hook, _ = mwl(inputs=inputs, labels=y, batch_shared_entities=shared, use_batch_shared_entities=shared_mode,
                            epoch=1, input_style_triple_or_prefix="right_and_left_prefix")
        (loss.sum() / nl).backward()                                          # openkge/trainer.py:217-234
        for o in opts:
            o.step()                                                          # :238-246
        return time.perf_counter() - t0, nm

    t_probe, _ = one_step(min(B, 64))
    per_row = t_probe / min(B, 64)
    b_fit = int(max(8, min(B, budget_s / max(steps + warmup, 1) / max(per_row, 1e-9))))
    for _ in range(max(warmup - 1, 0)):
        one_step(b_fit)
    total_t, total_m = 0.0, 0.0
    for _ in range(steps):
        dt, nm = one_step(b_fit)
        total_t += dt
        total_m += nm
    sample = (f"{steps} training steps of the unmodified reference (AddLossModule.forward + backward + OptimRegime.step, "
              f"openkge/trainer.py:48-113, 217-246) on {b_fit} prefix rows x {n_cols[-1]} candidates, D={wl['dim']}, dense fp32 "
              f"labels, pre-collated, {cores} torch threads")
    return total_m / 2.0 / total_t, total_t / steps * 1e3, cores, sample, b_fit


def eval_run(workloads, workload, steps, budget_s, seed=1, rows_per_step=32):
    """Reference filtered evaluation (forward + loss + compute_metrics, openkge/trainer.py:259-272) on `rows_per_step`
    prefix rows per step (the reference's own test batch size). Returns (ranked answers/s, ms/step, cores, sample)."""
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    _, RefDataset, _, _, _ = _import_reference()
    wl = workloads[workload]
    spec = S.SPECS[wl["spec"]]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, ev_idx, meta = S.build_indexes(spec, seed=seed, scale=wl.get("scale", 1.0))
    model, mwl, _ = _build_model(wl, meta, seed)
    model.eval()
    t_pre = 0.0
    if hasattr(model, "entity_token_ids"):                   # token models: cached encode of every entity
        t0 = time.perf_counter()
        with torch.no_grad():
            model.precompute_embeddings_from_tokens()        # cached encode of every entity (amortised over the split)
        t_pre = time.perf_counter() - t0
    rng = np.random.default_rng(seed + 11)
    total_t, total_q, done = 0.0, 0, 0
    with torch.no_grad():
        while done < steps and total_t < budget_s:
            batch = ev_idx.collate(rng.integers(0, len(ev_idx), rows_per_step))
            inputs, nl, nm, y, label_ids, fmask, shared = _dense_wire_batch(batch, spec.n_entities, False)
            t0 = time.perf_counter()
            loss, _, scores = mwl(inputs=inputs, labels=y, batch_shared_entities=shared, use_batch_shared_entities=False,
                                  epoch=1, input_style_triple_or_prefix="right_and_left_prefix")
            res = RefDataset.compute_metrics(fmask, label_ids, scores)
            total_t += time.perf_counter() - t0
            total_q += int(res["mrr"].count)
            done += 1
    sample = (f"{done} eval steps of the unmodified reference (forward + loss + compute_metrics, openkge/trainer.py:259-272) on "
              f"{rows_per_step} prefix rows x {spec.n_entities} candidates, D={wl['dim']}, dense labels + filter mask"
              + (f"; entity cache precomputed once in {t_pre:.1f} s, not counted" if t_pre else "") + f", {cores} torch threads")
    return total_q / max(total_t, 1e-9), total_t / max(done, 1) * 1e3, cores, sample
