"""N > 1 host logic of the entity-sharded path on CPU: two gloo ranks, each owning half of the entity rows,
must reproduce the single-device result (loss, post-step weights of every shard, rank counts bit-exact).
The compute engine is the oracle-backed stand-in of tests/oracle_engine.py (there is no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_knowledge_graph_embeddings_b200 import dataset as D
from open_knowledge_graph_embeddings_b200.sharded import EntityShardedLookupModel, restrict_csr, shard_bounds
from oracle import okge_oracle as O
from tests import oracle_engine
from tests.conftest import load_golden, params_of


def test_shard_bounds_cover_everything():
    for n, w in [(10, 3), (1000000, 8), (7, 8), (14541, 4)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_restrict_csr():
    c = D.CSRMatrix.from_lists([[0, 5, 9], [], [4, 5], [9]], 10)
    ptr, idx = restrict_csr(c.ptr, c.idx, 4, 9)
    assert ptr.tolist() == c.ptr.tolist() and idx.tolist() == [-1, 1, -1, 0, 1, -1]     # foreign columns are marked, not dropped


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batches(gold):
    t = lambda k: torch.from_numpy(gold[k]).view(-1, 1)
    out = {}
    for split in ("train", "eval"):
        po, sp = (t(f"{split}/po_rel"), t(f"{split}/po_obj")), (t(f"{split}/sp_subj"), t(f"{split}/sp_rel"))
        B = po[0].numel() + sp[0].numel()
        N = int(gold[f"{split}/normalizer_loss"]) // B
        labels = D.CSRMatrix(torch.from_numpy(gold[f"{split}/pos_ptr"]), torch.from_numpy(gold[f"{split}/pos_idx"]), (B, N))
        ans = filt = None
        if split == "eval":
            filt = D.CSRMatrix(torch.from_numpy(gold["eval/filt_ptr"]), torch.from_numpy(gold["eval/filt_idx"]), (B, N))
            ans = D.RankedAnswers(*(torch.from_numpy(gold[f"eval/{k}"]) for k in ("ans_row", "alt_ptr", "alt_idx")))
        out[split] = ([po, sp], B * N, float(labels.nnz), labels, ans, filt, None)
    return out


def _run_rank(rank, world, port, name, scorer, loss, smoothing, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    gold = load_golden(name)
    batches = _batches(gold)
    W = torch.from_numpy(gold["init/entity_embedding.weight"])
    N = W.size(0) - 2
    lo, hi = shard_bounds(N, world, rank)
    model = EntityShardedLookupModel.from_reference_state_dict(
        {k: torch.from_numpy(v) for k, v in params_of(gold, "init/").items()}, rank, world, "cpu", scorer=scorer, lr=0.3,
        eps=1e-8, weight_decay=1e-10, engine=oracle_engine,
        fused_update_max_rows=0 if scorer == "complex" else 3072)     # ComplEx case: the unfused (large-batch) update path
    assert (model.lo, model.hi) == (lo, hi) and torch.equal(model.E, W[2 + lo:2 + hi])
    # evaluation first, on identical (initial) weights: counts and true scores must then be bit-identical for any
    # number of shards; after a training step the weights differ in the last bits (summation order of dQ)
    true, greater, equal = model.eval_counts(batches["eval"])
    loss_sum = model.train_step(batches["train"], smoothing=smoothing, loss=loss)
    torch.save({"loss": float(loss_sum), "E": model.E, "R": model.R, "lo": lo, "hi": hi, "true": true,
                "greater": greater, "equal": equal}, f"{result_path}.{world}.{rank}")
    torch.save(model.state_dict_shard(), f"{result_path}.ckpt.{world}.{rank}")
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,scorer,loss,smoothing", [
    ("lookup_distmult_bce", "distmult", "bce", 0.0),
    ("lookup_complex_bce_smooth", "complex", "bce", 0.1),
])
def test_two_gloo_ranks_match_single_device_and_reference(tmp_path, name, scorer, loss, smoothing):
    path = str(tmp_path / "res")
    _run_rank(0, 1, _free_port(), name, scorer, loss, smoothing, path)
    mp.spawn(_run_rank, args=(2, _free_port(), name, scorer, loss, smoothing, path), nprocs=2, join=True)
    one = torch.load(f"{path}.1.0")
    two = [torch.load(f"{path}.2.{r}") for r in range(2)]
    gold = load_golden(name)
    # loss of the reference
    assert one["loss"] == pytest.approx(float(gold["train/loss_sum"]), rel=1e-5)
    assert two[0]["loss"] == pytest.approx(one["loss"], rel=1e-9) and two[1]["loss"] == pytest.approx(one["loss"], rel=1e-9)
    # post-step weights: shards concatenate to the single-device table, which matches the reference's step
    E2 = torch.cat([two[0]["E"], two[1]["E"]])
    np.testing.assert_allclose(E2.numpy(), one["E"].numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(one["E"].numpy(), gold["step1/entity_embedding.weight"][2:], rtol=2e-4, atol=2e-5)
    for r in range(2):
        np.testing.assert_allclose(two[r]["R"].numpy(), one["R"].numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(one["R"].numpy()[2:], gold["step1/relation_embedding.weight"][2:], rtol=2e-4, atol=2e-5)
    # rank counts: identical on both ranks, equal to the single-device counts, and to the oracle on the same scores
    for k in ("greater", "equal"):
        assert torch.equal(two[0][k], two[1][k]) and torch.equal(two[0][k], one[k]), k
    assert torch.equal(two[0]["true"], one["true"]) and torch.equal(two[1]["true"], one["true"])
    W0 = gold["init/entity_embedding.weight"]
    om = O.OracleModel("lookup", scorer, params_of(gold, "init/"))
    scores = om.scores(gold["eval/po_rel"], gold["eval/po_obj"], gold["eval/sp_subj"], gold["eval/sp_rel"]).astype(np.float64)
    _, og, oe = O.rank_counts(scores.astype(np.float32), gold["eval/ans_row"], gold["eval/alt_ptr"], gold["eval/alt_idx"],
                              gold["eval/filt_ptr"], gold["eval/filt_idx"])
    ranks = one["greater"].numpy() + one["equal"].numpy() // 2
    assert np.abs(ranks - (og + oe // 2)).max() <= 1      # fp64 engine vs fp32 oracle scores: near-ties may flip


def _resume_rank(rank, world, port, name, ckpt_path, result_path):
    """Second training step from a checkpoint that was re-partitioned from `ckpt_path` (any number of source ranks)."""
    from open_knowledge_graph_embeddings_b200.sharded import merge_lookup_shards
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    gold = load_golden(name)
    state, sums, steps = merge_lookup_shards([torch.load(p) for p in ckpt_path])
    model = EntityShardedLookupModel.from_reference_state_dict(state, rank, world, "cpu", scorer="distmult", adagrad_sums=sums,
                                                               training_steps=steps, lr=0.3, eps=1e-8, weight_decay=1e-10,
                                                               engine=oracle_engine)
    loss_sum = model.train_step(_batches(gold)["train"])
    torch.save({"loss": float(loss_sum), "shard": model.state_dict_shard()}, f"{result_path}.{world}.{rank}")
    if world > 1:
        dist.destroy_process_group()


def test_sharded_checkpoint_roundtrip_and_repartition(tmp_path):
    """Per-rank shard files merge into the reference's state dict (keys, shapes, PAD / UNK rows decayed like the dense
    reference step), load into the single-device model class with strict=True, and re-partition onto a different number
    of ranks: 2 ranks -> merge -> 1 rank and 1 rank -> merge -> 2 ranks continue to the same second step."""
    from open_knowledge_graph_embeddings_b200.dataset import EntityRelationDatasetMeta
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.sharded import merge_lookup_shards
    name = "lookup_distmult_bce"
    path = str(tmp_path / "res")
    _run_rank(0, 1, _free_port(), name, "distmult", "bce", 0.0, path)
    mp.spawn(_run_rank, args=(2, _free_port(), name, "distmult", "bce", 0.0, path), nprocs=2, join=True)
    gold = load_golden(name)
    ck1 = [f"{path}.ckpt.1.0"]
    ck2 = [f"{path}.ckpt.2.0", f"{path}.ckpt.2.1"]
    s1, g1, t1 = merge_lookup_shards([torch.load(p) for p in ck1])
    s2, g2, t2 = merge_lookup_shards([torch.load(p) for p in reversed(ck2)])          # any order
    assert t1 == t2 == 1
    for k in ("entity_embedding.weight", "relation_embedding.weight"):
        assert s2[k].shape == tuple(gold["step1/" + k].shape)
        np.testing.assert_allclose(s2[k].numpy(), s1[k].numpy(), rtol=1e-5, atol=1e-6)
        # every row, including PAD / UNK (rows 0, 1: no gradient, decayed by the dense update), matches the reference's step
        np.testing.assert_allclose(s1[k].numpy(), gold["step1/" + k], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(g1[k].numpy(), gold[f"optstate/{k}/sum"], rtol=2e-3, atol=1e-12)
    meta = EntityRelationDatasetMeta(entities_size=s2["entity_embedding.weight"].shape[0],
                                     relations_size=s2["relation_embedding.weight"].shape[0])
    single = Models.LookupDistmultRelationModel(entity_slot_size=s2["entity_embedding.weight"].shape[1], init_std=0.1,
                                                train_data=meta)
    res = single.load_state_dict(s2, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(ValueError):
        merge_lookup_shards([torch.load(ck2[0])])                                      # a rank is missing
    # continue training from the re-partitioned checkpoints
    out = str(tmp_path / "resume")
    _resume_rank(0, 1, _free_port(), name, ck2, out + ".from2")
    mp.spawn(_resume_rank, args=(2, _free_port(), name, ck1, out + ".from1"), nprocs=2, join=True)
    a = torch.load(out + ".from2.1.0")
    b = [torch.load(f"{out}.from1.2.{r}") for r in range(2)]
    assert b[0]["loss"] == pytest.approx(a["loss"], rel=1e-6) and b[1]["loss"] == pytest.approx(a["loss"], rel=1e-6)
    sa, _, ta = merge_lookup_shards([a["shard"]])
    sb, _, tb = merge_lookup_shards([x["shard"] for x in b])
    assert ta == tb == 2
    for k in sa:
        np.testing.assert_allclose(sb[k].numpy(), sa[k].numpy(), rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------------------------------------
# token model: candidates partitioned, token tables replicated
# ---------------------------------------------------------------------------------------------

def _run_rank_unigram(rank, world, port, name, pool, result_path):
    from open_knowledge_graph_embeddings_b200.sharded import CandidateShardedUnigramModel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    gold = load_golden(name)
    batches = _batches(gold)
    params = {k: torch.from_numpy(v).clone() for k, v in params_of(gold, "init/").items()}
    N = params["entity_token_ids"].size(0) - 2
    model = CandidateShardedUnigramModel(params, N, rank, world, scorer="complex", pool=pool, lr=0.3, eps=1e-8,
                                         weight_decay=1e-10, engine=oracle_engine)
    true, greater, equal = model.eval_counts(batches["eval"])
    loss_sum = model.train_step(batches["train"])
    out = {"loss": float(loss_sum), "true": true, "greater": greater, "equal": equal}
    out.update({k: v for k, v in model.p.items() if v.dtype.is_floating_point})
    torch.save(out, f"{result_path}.{world}.{rank}")
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,pool", [("unigram_complex_sum_bn_bce", "sum"), ("unigram_complex_mean_bce", "mean")])
def test_two_gloo_ranks_token_model_match_single_device_and_reference(tmp_path, name, pool):
    """Candidate-sharded UnigramPooling model: 2 gloo ranks == 1 rank == the reference's golden training step
    (loss, post-step token tables, batch-norm parameters and running statistics) and its filtered rank counts."""
    path = str(tmp_path / "res")
    _run_rank_unigram(0, 1, _free_port(), name, pool, path)
    mp.spawn(_run_rank_unigram, args=(2, _free_port(), name, pool, path), nprocs=2, join=True)
    one = torch.load(f"{path}.1.0")
    two = [torch.load(f"{path}.2.{r}") for r in range(2)]
    gold = load_golden(name)
    assert one["loss"] == pytest.approx(float(gold["train/loss_sum"]), rel=1e-5)
    for r in range(2):
        assert two[r]["loss"] == pytest.approx(one["loss"], rel=1e-9)
    keys = [k for k in one if k not in ("loss", "true", "greater", "equal")]
    assert "entity_embedding.weight" in keys and "relation_embedding.weight" in keys
    for k in keys:
        # The first Adagrad step is sign-like, lr * g / (|g| + 1e-8): elements whose gradient is pure cancellation noise
        # (e.g. the EOS token row under batch norm: every mention contains it, its gradient is ~1e-9) are not comparable;
        # everything with a gradient above the noise floor must agree.
        solid = np.ones(one[k].shape, bool)
        if ("grad/" + k) in gold:
            g = np.abs(gold["grad/" + k])
            solid = g > 1e-4 * g.max()
            assert solid.mean() > 0.5, k
        for r in range(2):                                   # replicated state stays identical on both ranks
            np.testing.assert_allclose(two[r][k].numpy()[solid], one[k].numpy()[solid], rtol=1e-5, atol=1e-6, err_msg=k)
            assert torch.equal(two[0][k], two[1][k]) or np.allclose(two[0][k].numpy(), two[1][k].numpy(), rtol=0, atol=0), k
        np.testing.assert_allclose(one[k].numpy()[solid], gold["step1/" + k][solid], rtol=3e-4, atol=3e-5, err_msg=k)
    for k in ("greater", "equal"):
        assert torch.equal(two[0][k], two[1][k]) and torch.equal(two[0][k], one[k]), k
    assert torch.equal(two[0]["true"], one["true"]) and torch.equal(two[1]["true"], one["true"])
    # the evaluation ran on the initial weights (the golden's eval block is post-step): compare with the oracle model
    om = O.OracleModel("unigram", "complex", params_of(gold, "init/"), pool=pool, batchnorm="bn" in name)
    scores = om.scores(gold["eval/po_rel"], gold["eval/po_obj"], gold["eval/sp_subj"], gold["eval/sp_rel"], training=False)
    _, og, oe = O.rank_counts(scores.astype(np.float32), gold["eval/ans_row"], gold["eval/alt_ptr"], gold["eval/alt_idx"],
                              gold["eval/filt_ptr"], gold["eval/filt_idx"])
    ranks = one["greater"].numpy() + one["equal"].numpy() // 2
    assert np.abs(ranks - (og + oe // 2)).max() <= 1          # fp64 engine vs fp32 oracle scores: near-ties may flip
