"""Pins the oracle (oracle/okge_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, written by tests/golden/make_golden.py). CPU only.

Tolerances: float quantities agree with the reference's fp32 torch results to fp32 round-off
(rtol 2e-5 / atol 2e-6: summation order differs); rank counts and every integer tensor are exact.
"""
import numpy as np
import pytest

from oracle import okge_oracle as O
from tests.conftest import load_golden, params_of

RTOL, ATOL = 2e-5, 2e-6


def test_pack_unpack_kat(kats):
    assert O.pack_list_of_lists([[5], [6, 7], [8]]) == kats["pack/a"].tolist() == [5, 6, 8, 9, 0, 5, 6, 7, 8]
    assert O.pack_list_of_lists([9, 10]) == kats["pack/b"].tolist() == [4, 5, 6, 0, 9, 10]
    lol, flat = O.unpack_list_of_lists(kats["pack/a"])
    assert lol == [[5], [6, 7], [8]] and flat == kats["unpack/a_flat"].tolist()
    assert O.unpack_list_of_lists(O.pack_list_of_lists([])) == ([], [])


def test_compute_metrics_kat(kats):
    t, g, e = O.rank_counts(kats["kat/scores"], kats["kat/ans_row"], kats["kat/alt_ptr"], kats["kat/alt_idx"],
                            kats["kat/filt_ptr"], kats["kat/filt_idx"])
    assert np.array_equal(t, kats["kat/true"])
    assert g.tolist() == kats["kat/greater"].tolist() == [1, 1, 0]
    assert e.tolist() == kats["kat/equal"].tolist() == [3, 0, 7]
    assert (g + e // 2).tolist() == [2, 1, 3]
    m = O.compute_metrics(kats["kat/scores"], kats["kat/ans_row"], kats["kat/alt_ptr"], kats["kat/alt_idx"],
                          kats["kat/filt_ptr"], kats["kat/filt_idx"])
    for k in ("mrr", "mr", "h1", "h3", "h10", "h50"):
        assert m[k].avg == pytest.approx(kats[f"kat/metric/{k}"][0], rel=1e-6, abs=1e-9)
        assert m[k].count == int(kats[f"kat/metric/{k}"][1]) == 3
    assert m["mrr"].avg == pytest.approx(0.361111, abs=1e-6)


@pytest.mark.parametrize("split,training", [("train", True), ("valid", False)])
def test_collate_matches_reference(kats, split, training):
    sizes = kats["meta/sizes"]
    rows = kats[f"data/{split}/seen_prefixes"][kats[f"collate/{split}/sampler"]]
    c = O.collate_full(rows, kats[f"data/{split}/seen_entities"], kats[f"data/{split}/all_splits_entities"],
                       entity_vocab_size=int(sizes[0]), entity_vocab_offset=2, is_training_data=training)
    for k in ("po", "sp", "pos_ptr", "pos_idx"):
        assert np.array_equal(c[k], kats[f"collate/{split}/{k}"]), k
    assert [c["normalizer_loss"], c["normalizer_metric"]] == kats[f"collate/{split}/normalizers"].tolist()
    if not training:
        for k in ("filt_ptr", "filt_idx", "ans_row", "alt_ptr", "alt_idx"):
            assert np.array_equal(c[k], kats[f"collate/{split}/{k}"]), k


def _oracle_model(case, params):
    kind, scorer, _, _, pool, bn, _ = case
    return O.OracleModel(kind, scorer, params, pool=pool, batchnorm=bn)


def test_train_step_matches_reference(model_case):
    name, case, gold = model_case
    _, _, loss, smoothing, _, _, optimizer = case
    m = _oracle_model(case, params_of(gold, "init/"))
    scores, loss_sum, grads = m.loss_and_grads(gold["train/po_rel"], gold["train/po_obj"], gold["train/sp_subj"],
                                               gold["train/sp_rel"], gold["train/pos_ptr"], gold["train/pos_idx"],
                                               loss=loss, smoothing=smoothing)
    # the LSTM recurrence runs in fp32 in the reference and in fp64 here: ten steps of round-off, scaled up by batch norm
    atol = 2e-5 if case[0] == "lstm" else ATOL
    np.testing.assert_allclose(scores, gold["train/scores"], rtol=RTOL, atol=atol)
    assert loss_sum == pytest.approx(float(gold["train/loss_sum"]), rel=2e-6)
    assert scores.shape[0] * scores.shape[1] == int(gold["train/normalizer_loss"])
    for k, g in params_of(gold, "grad/").items():
        scale = np.abs(g).max() + 1e-30
        np.testing.assert_allclose(grads[k], g, rtol=1e-4, atol=2e-5 * scale, err_msg=k)
    # optimizer step with the hyper-parameters the reference EFFECTIVELY used
    lr, eps, wd = float(gold["opt/lr"]), float(gold["opt/eps"]), float(gold["opt/weight_decay"])
    assert eps == 1e-8
    for k, g in params_of(gold, "grad/").items():
        p0 = gold["init/" + k]
        if optimizer == "adagrad":
            p1, st = O.adagrad_step(p0, g, np.zeros_like(p0), lr, eps, wd)
            np.testing.assert_allclose(st, gold[f"optstate/{k}/sum"], rtol=1e-6, atol=1e-12)
        else:
            p1, m1, v1 = O.adam_step(p0, g, np.zeros_like(p0), np.zeros_like(p0), lr, 1, eps=eps, weight_decay=wd)
            np.testing.assert_allclose(m1, gold[f"optstate/{k}/exp_avg"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(p1, gold["step1/" + k], rtol=2e-6, atol=1e-7, err_msg=k)


def test_eval_matches_reference(model_case):
    name, case, gold = model_case
    _, _, loss, smoothing, _, _, _ = case
    m = _oracle_model(case, params_of(gold, "step1/"))
    scores = m.scores(gold["eval/po_rel"], gold["eval/po_obj"], gold["eval/sp_subj"], gold["eval/sp_rel"],
                      training=False)
    np.testing.assert_allclose(scores, gold["eval/scores"], rtol=5e-5, atol=5e-6)
    # rank counts: bit-exact on IDENTICAL scores (the reference's own predictions)
    args = (gold["eval/ans_row"], gold["eval/alt_ptr"], gold["eval/alt_idx"], gold["eval/filt_ptr"], gold["eval/filt_idx"])
    t, g, e = O.rank_counts(gold["eval/scores"], *args)
    assert np.array_equal(t, gold["eval/true_score"])
    assert np.array_equal(g, gold["eval/greater"])
    assert np.array_equal(e, gold["eval/equal"])
    met = O.compute_metrics(gold["eval/scores"], *args)
    for k in ("mrr", "mr", "h1", "h3", "h10", "h50"):
        assert met[k].avg == pytest.approx(gold[f"eval/metric/{k}"][0], rel=1e-6, abs=1e-9), k
        assert met[k].count == int(gold[f"eval/metric/{k}"][1])
    # eval loss
    y = O.dense_labels(gold["eval/pos_ptr"], gold["eval/pos_idx"], scores.shape[1])
    if loss == "bce":
        ls = O.bce_with_logits_sum(gold["eval/scores"], O.smooth_labels(y, smoothing))
    else:
        ls = O.kl_log_softmax_sum(gold["eval/scores"], y)
    assert ls == pytest.approx(float(gold["eval/loss_sum"]), rel=3e-6)


def test_fold_equals_reference_form():
    rng = np.random.default_rng(0)
    a, b, e = (rng.standard_normal((5, 12)).astype(np.float32) for _ in range(3))
    E = rng.standard_normal((40, 12)).astype(np.float32)
    np.testing.assert_allclose(O.fold_query(O.FOLD_COMPLEX_SP, a, b) @ E.T, O.complex_prefix_score(a, b, E, True), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(O.fold_query(O.FOLD_COMPLEX_PO, a, b) @ E.T, O.complex_prefix_score(E, b, a, False), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(O.fold_query(O.FOLD_DISTMULT, a, b) @ E.T, O.distmult_prefix_score(a, b, E, True), rtol=1e-4, atol=1e-5)


# ---- the torch-CPU port used as the timed CPU baseline is pinned to the same golden vectors ----

def _port(case, params):
    from oracle import torch_cpu_port as P
    kind, scorer, _, _, pool, bn, _ = case
    return P, P.PortModel(kind, scorer, params, pool=pool, batchnorm=bn)


def _pairs(gold, split):
    import torch
    t = lambda k: torch.from_numpy(gold[f"{split}/{k}"]).view(-1, 1)
    return (t("po_rel"), t("po_obj")), (t("sp_subj"), t("sp_rel"))


def test_torch_cpu_port_train_step(model_case):
    import torch
    name, case, gold = model_case
    _, _, loss, smoothing, _, _, optimizer = case
    P, m = _port(case, params_of(gold, "init/"))
    po, sp = _pairs(gold, "train")
    N = int(gold["train/normalizer_loss"]) // (po[0].numel() + sp[0].numel())
    y = P.dense_labels(gold["train/pos_ptr"], gold["train/pos_idx"], N)
    opt = (P.make_adagrad(m, float(gold["opt/lr"]), float(gold["opt/weight_decay"])) if optimizer == "adagrad"
           else torch.optim.Adam(m.parameters(), lr=float(gold["opt/lr"])))
    loss_v, scores = P.train_step(m, opt, po, sp, y, loss, smoothing)
    np.testing.assert_allclose(scores.detach().numpy(), gold["train/scores"], rtol=RTOL, atol=ATOL)
    assert loss_v == pytest.approx(float(gold["train/loss_sum"]), rel=2e-6)
    for k in params_of(gold, "grad/"):
        np.testing.assert_allclose(dict(m.named_parameters())[k].detach().numpy(), gold["step1/" + k], rtol=1e-5, atol=1e-6,
                                   err_msg=k)


def test_torch_cpu_port_eval(model_case):
    import torch
    name, case, gold = model_case
    P, m = _port(case, params_of(gold, "step1/"))
    po, sp = _pairs(gold, "eval")
    B = po[0].numel() + sp[0].numel()
    N = int(gold["eval/normalizer_loss"]) // B
    y = P.dense_labels(gold["eval/pos_ptr"], gold["eval/pos_idx"], N)
    filt = P.dense_labels(gold["eval/filt_ptr"], gold["eval/filt_idx"], N).bool()
    label_ids = [[] for _ in range(B)]
    for j, b in enumerate(gold["eval/ans_row"]):
        label_ids[b].append(torch.from_numpy(gold["eval/alt_idx"][gold["eval/alt_ptr"][j]:gold["eval/alt_ptr"][j + 1]]))
    _, mrr_sum, count, ranks = P.eval_step(m, po, sp, y, filt, label_ids, case[2])
    assert count == int(gold["eval/metric/mrr"][1])
    assert mrr_sum / count == pytest.approx(gold["eval/metric/mrr"][0], rel=1e-5)


@pytest.mark.parametrize("name,kind,bn", [("traj_lookup_complex", "lookup", False), ("traj_unigram_bn", "unigram", True)])
def test_oracle_follows_the_reference_training_trajectory(name, kind, bn):
    """30 optimizer steps of the unmodified reference (loss of every step) + the filtered evaluation of the whole
    validation split, replayed by the numpy oracle from the same initial weights on the same batches."""
    gold, kat = load_golden(name), load_golden("kats")
    m = O.OracleModel(kind, "complex", params_of(gold, "init/"), pool="sum", batchnorm=bn)
    sums = {k: np.zeros_like(v) for k, v in m.p.items()
            if k.endswith("embedding.weight") or k.endswith("batchnorm.weight") or k.endswith("batchnorm.bias")}
    n_ent = int(kat["meta/sizes"][0])
    losses = []
    for rows in gold["traj/rows"]:
        b = O.collate_full(kat["data/train/seen_prefixes"][rows], kat["data/train/seen_entities"],
                           kat["data/train/all_splits_entities"], n_ent, 2, True)
        _, loss_sum, grads = m.loss_and_grads(b["po"][:, 0], b["po"][:, 1], b["sp"][:, 0], b["sp"][:, 1], b["pos_ptr"],
                                              b["pos_idx"])
        losses.append(loss_sum / b["normalizer_loss"])
        for k in sums:
            m.p[k], sums[k] = O.adagrad_step(m.p[k], grads[k], sums[k], 0.3, 1e-8, 1e-10)
    np.testing.assert_allclose(losses, gold["traj/loss"], rtol=2e-4)
    # Adagrad's first steps are sign-like (g / (|g| + eps)): an element whose gradient is at the round-off level can take a
    # different step in fp64 (here) than in fp32 (reference); everything else lands on the reference's final weights
    for k in sums:
        assert np.isclose(m.p[k], gold["final/" + k], rtol=2e-3, atol=2e-4).mean() > 0.9, k
    # evaluation of the whole validation split in file order
    P = kat["data/valid/seen_prefixes"]
    bs = gold["traj/rows"].shape[1]
    ranks, rows_of = [], []
    for lo in range(0, len(P), bs):
        b = O.collate_full(P[lo:lo + bs], kat["data/valid/seen_entities"], kat["data/valid/all_splits_entities"], n_ent, 2, False)
        scores = m.scores(b["po"][:, 0], b["po"][:, 1], b["sp"][:, 0], b["sp"][:, 1], training=False)
        _, g, e = O.rank_counts(scores, b["ans_row"], b["alt_ptr"], b["alt_idx"], b["filt_ptr"], b["filt_idx"])
        ranks.append(g + e // 2)
        rows_of.append(b["ans_row"] + lo)
    met = O.metrics_from_ranks(np.concatenate(ranks), np.concatenate(rows_of))
    n = int(gold["eval/metric/mrr"][1])
    assert met["mrr"].count == n
    assert abs(met["mrr"].avg - gold["eval/metric/mrr"][0]) < 1e-3
    for k in ("h1", "h3", "h10", "h50"):
        assert abs(met[k].avg - gold[f"eval/metric/{k}"][0]) <= 1.0 / n + 1e-9, k
