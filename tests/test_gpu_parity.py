"""GPU parity tests (run on the B200 with ``-m gpu``): the native CUDA path, called through the C ABI,
against (a) the numpy oracle on seeded inputs, (b) the golden vectors of the unmodified reference,
(c) size-independent properties at larger sizes.

Tolerances (north_star: rel 1e-3 on scores and loss, bit-exact ranks and filtered counts):
  * integer / index work (rank counts, gathers): bit-exact;
  * fp32 CUDA-core kernels (pooling, fold, optimizers): 1e-6 relative to the operand scale;
  * tensor-core scores (FP16 inputs = 10-bit mantissa rounded to nearest, FP32 accumulate):
    |ds| <= 1e-3 * ||q|| * ||e|| (norm-wise); split precision (evaluation, hi + lo planes): <= 4e-6;
  * loss: 1e-3 relative;   gradients: 2e-3 of the largest gradient entry.
"""
import os

import numpy as np
import pytest
import torch

from oracle import okge_oracle as O
from tests.conftest import MODEL_CASES, load_golden, params_of

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-3
LOSS_RTOL = 1e-3
GRAD_TOL = 2e-3


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from open_knowledge_graph_embeddings_b200 import kernels
    return kernels


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def normwise(scores, ref, q, e):
    bound = np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(e, axis=1)[None, :]
    return float((np.abs(scores - ref) / np.maximum(bound, 1e-30)).max())


def random_csr(rng, B, N, max_per_row, empty_rows=()):
    rows = [sorted(set(rng.integers(0, N, rng.integers(1, max_per_row + 1)).tolist())) for _ in range(B)]
    for r in empty_rows:
        rows[r] = []
    ptr = np.zeros(B + 1, np.int32)
    ptr[1:] = np.cumsum([len(r) for r in rows])
    return ptr, np.concatenate([np.asarray(r, np.int64) for r in rows]).astype(np.int32)


# ---------------------------------------------------------------------------------------------
# kernels vs oracle
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
@pytest.mark.parametrize("D", [64, 200])
def test_gather_pool_fwd_bwd(K, mode, D):
    rng = np.random.default_rng(1)
    V, R, L, n = 300, 150, 10, 97
    W = rng.standard_normal((V, D)).astype(np.float32)
    rows = rng.integers(0, V, (R, L))
    rows[:, 5:] = 0                      # PAD slots, gathered like every other slot
    rows[7] = 0                          # an all-PAD row: mean divides by 1e-12
    ids = rng.integers(0, R, n).astype(np.int32)
    out = K.gather_pool_fwd(dev(W), dev(rows, torch.int32), dev(ids), mode).cpu().numpy()
    ref = O.unigram_pool_encode(W, rows, ids, mode)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    # all-rows variant (ids = None, id_start)
    out2 = K.gather_pool_fwd(dev(W), dev(rows, torch.int32), None, mode, id_start=2).cpu().numpy()
    np.testing.assert_allclose(out2, O.unigram_pool_encode(W, rows, np.arange(2, R), mode), rtol=1e-5,
                               atol=1e-5 * np.abs(ref).max())
    g = rng.standard_normal((n, D)).astype(np.float32)
    gw = torch.zeros(V, D, device="cuda")
    K.gather_pool_bwd(dev(g), dev(W), dev(rows, torch.int32), dev(ids), mode, gw)
    gref = O.unigram_pool_backward(g, W, rows, ids, mode)
    assert np.all(gw.cpu().numpy()[0] == 0), "padding_idx row must not receive gradient"
    np.testing.assert_allclose(gw.cpu().numpy(), gref, rtol=1e-4, atol=1e-5 * np.abs(gref).max())


@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
def test_gather_pool_compact_slot_gradient(K, mode):
    """GatherPool with the compact (slot) token gradient: materialised it equals the dense backward, applied through
    okge_adagrad_slot_table it equals the dense Adagrad step on that gradient over the WHOLE table (rows without a slot
    still take their weight-decay step), and the slot map is handed back clean."""
    from open_knowledge_graph_embeddings_b200 import functional as Fn
    rng = np.random.default_rng(7)
    V, D, L, rows_n, n = 6000, 64, 10, 900, 120
    W = torch.nn.Parameter(dev((0.3 * rng.standard_normal((V, D))).astype(np.float32)))
    tok = rng.integers(4, V, size=(rows_n, L))
    tok[:, 0] = 2                                                     # BOS everywhere: the privatised hot rows
    tok[np.arange(L)[None, :] >= rng.integers(2, L + 1, size=(rows_n, 1))] = 0
    id_rows = dev(tok.astype(np.int32))
    ids = dev(rng.integers(0, rows_n, n).astype(np.int32))
    g = dev(rng.standard_normal((n, D)).astype(np.float32))
    Fn.GatherPool.apply(W, id_rows, ids, mode).backward(g)
    dense = W.grad.clone()
    W.grad = None
    W._okge_slot_update = True
    Fn.GatherPool.apply(W, id_rows, ids, mode).backward(g)
    d = W._okge_deferred
    assert W.grad is None and isinstance(d, Fn.SlotTableGrad)
    p1, s1 = W.detach().clone(), torch.rand_like(W) * 0.01
    p2, s2 = p1.clone(), s1.clone()
    K.adagrad_dense(p1, dense, s1, 0.1, 1e-8, 1e-4)
    d.adagrad_step(torch.nn.Parameter(p2), s2, 0.1, 1e-8, 1e-4)
    # The two gradients are the same sums taken in different (run-dependent) orders of float atomics: the BOS row adds up
    # 120 terms of magnitude ~1 per column, i.e. |dg| up to ~1e-5, whatever is left of g after cancellation. That enters
    # G as 2 g dg and the parameter as lr dg / sqrt(G) with sqrt(G) >= 0.03 here: absolute bounds, not relative ones.
    np.testing.assert_allclose(p2.cpu().numpy(), p1.cpu().numpy(), rtol=1e-5, atol=5e-5)
    np.testing.assert_allclose(s2.cpu().numpy(), s1.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert torch.all(W._okge_slot_map == -1)
    # materialised form (the Adam / clipping route) and the discard of an unconsumed gradient
    W._okge_deferred = None
    Fn.GatherPool.apply(W, id_rows, ids, mode).backward(g)
    np.testing.assert_allclose(W._okge_deferred.materialize().cpu().numpy(), dense.cpu().numpy(), rtol=1e-5, atol=2e-5)
    assert torch.all(W._okge_slot_map == -1)
    W._okge_deferred = None
    Fn.GatherPool.apply(W, id_rows, ids, mode).backward(g)
    W._okge_deferred.discard()
    assert torch.all(W._okge_slot_map == -1)


def test_gather_pool_empty_and_errors(K):
    W = torch.zeros(10, 8, device="cuda")
    rows = torch.zeros(4, 10, dtype=torch.int32, device="cuda")
    assert K.gather_pool_fwd(W, rows, torch.zeros(0, dtype=torch.int32, device="cuda"), "sum").shape == (0, 8)
    from open_knowledge_graph_embeddings_b200._capi import OkgeNativeError
    with pytest.raises(OkgeNativeError):
        K.gather_pool_fwd(torch.zeros(10, 6, device="cuda"), rows, None, "sum")      # D % 4 != 0
    with pytest.raises(OkgeNativeError):
        K.gather_rows(torch.zeros(4, 8), torch.zeros(1, dtype=torch.int32))          # CPU tensor: no fallback


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_fold_query_fwd_bwd(K, kind):
    rng = np.random.default_rng(2)
    a, b, g = (rng.standard_normal((33, 24)).astype(np.float32) for _ in range(3))
    q = K.fold_query(kind, dev(a), dev(b)).cpu().numpy()
    np.testing.assert_allclose(q, O.fold_query(kind, a, b), rtol=1e-6, atol=1e-7)     # fp32; quantized where it is used
    ga, gb = K.fold_query_bwd(kind, dev(a), dev(b), dev(g))
    ra, rb = O.fold_query_backward(kind, a, b, g)
    np.testing.assert_allclose(ga.cpu().numpy(), ra, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gb.cpu().numpy(), rb, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("B,N,D", [(1, 1, 4), (5, 7, 8), (128, 256, 32), (129, 257, 36), (300, 1000, 200), (64, 5000, 512),
                                   (256, 40000, 64), (400, 20001, 36)])   # the last two: several waves of the persistent grid
def test_score_store_vs_oracle(K, B, N, D):
    """q E^T on the tcgen05 kernel vs the reference's own 4-product ComplEx form in float64."""
    rng = np.random.default_rng(B * 31 + N)
    s, r = (rng.standard_normal((B, D)).astype(np.float32) for _ in range(2))
    E = rng.standard_normal((N, D)).astype(np.float32)
    q = O.fold_query(O.FOLD_COMPLEX_SP, s, r)
    ref = O.complex_prefix_score(s.astype(np.float64), r.astype(np.float64), E.astype(np.float64), sp=True)
    out = K.score_store(dev(q), dev(E)).cpu().numpy()
    assert out.shape == (B, N)
    assert normwise(out, ref, q, E) < SCORE_TOL


@pytest.mark.parametrize("B,N,D", [(5, 7, 8), (129, 257, 36), (300, 1000, 200), (64, 5000, 512)])
def test_score_store_split_precision(K, B, N, D):
    """Three-term split-precision product (hi + lo fp16 planes): fp32-grade scores for the evaluation path."""
    rng = np.random.default_rng(B + N)
    q = rng.standard_normal((B, D)).astype(np.float32)
    E = (rng.standard_normal((N, D)) * np.exp(rng.standard_normal((N, 1)))).astype(np.float32)   # rows of mixed norms
    ref = q.astype(np.float64) @ E.astype(np.float64).T
    q16, e16 = K.quantize(dev(q), split=True), K.quantize(dev(E), split=True)
    out = K.score_store(q16, e16, split=True).cpu().numpy()
    assert normwise(out, ref, q, E) < 4e-6
    single = K.score_store(q16, e16).cpu().numpy()                # the same operands without their lo planes
    assert 1e-5 < normwise(single, ref, q, E) < SCORE_TOL
    # the operand represents the matrix to ~2^-22 of its largest element; hi alone to 2^-11 of each element
    assert np.abs(e16.dense().cpu().numpy() - E).max() <= 2.0 ** -21 * np.abs(E).max()
    assert np.abs(e16.without_lo().dense().cpu().numpy() - E).max() <= 2.0 ** -11 * np.abs(E).max()


def test_quantize_scales_and_edge_cases(K):
    """okge_f16_absmax + okge_f16_quantize: power-of-two scale with the largest element in [128, 256); tiny and huge
    tensors keep their relative precision; zeros and odd shapes (unaligned width, strided rows) are handled."""
    rng = np.random.default_rng(12)
    for mag in (1e-30, 1e-12, 1e-3, 1.0, 1e6, 1e30):
        x = (mag * rng.standard_normal((37, 50))).astype(np.float32)
        op = K.quantize(dev(x))
        inv = float(op.inv_scale.item())
        assert np.log2(inv) == np.round(np.log2(inv))                                  # exact power of two
        hi_max = float(op.hi[:, :50].float().abs().max())
        assert 128 <= hi_max < 256.5
        assert np.abs(op.dense().cpu().numpy() - x).max() <= 2.0 ** -11 * np.abs(x).max()
    z = K.quantize(torch.zeros(5, 16, device="cuda"))
    assert float(z.inv_scale.item()) == 1.0 and float(z.hi.float().abs().max()) == 0.0
    big = dev(rng.standard_normal((64, 100)).astype(np.float32))
    view = big[:, 4:84]                                                                # row pitch 100, offset 16 B
    op = K.quantize(view)
    assert np.abs(op.dense().cpu().numpy() - view.cpu().numpy()).max() <= 2.0 ** -11 * float(view.abs().max())
    fixed = K.quantize(view, fixed_scale=4.0)
    assert float(fixed.inv_scale.item()) == 0.25
    assert torch.equal(fixed.hi[:, :80], (view * 4.0).to(torch.float16))
    g = K.gather_rows_f16(K.quantize(big, split=True), dev(np.array([3, 3, 63, 0], np.int32)))
    assert g.shape == (4, 100) and np.allclose(g.dense().cpu().numpy(), big[[3, 3, 63, 0]].cpu().numpy(), rtol=0, atol=1e-5)


def test_gemm_split_k(K):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((70, 9000)).astype(np.float32)
    b = rng.standard_normal((40, 9000)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    for splits in (1, 5):
        out = K.gemm_nt(dev(a), dev(b), alpha=0.5, splits=splits).cpu().numpy()        # fp32 operands: TF32 contraction
        assert normwise(out, 0.5 * ref, a, b) < SCORE_TOL
    xp = K.Panels.from_dense(dev(a[:, :123]))
    assert xp.shape == (70, 123) and torch.all(xp.data[-1, :, 123 % 64:] == 0)         # zero tail of the last panel
    # panel operands through the tensor-core kernel (fp16)
    ap, bp = K.Panels.from_dense(dev(a)), K.Panels.from_dense(dev(b))
    for splits in (1, 4):
        out = K.gemm_nt(ap, bp, splits=splits).cpu().numpy()
        assert normwise(out, ref, a, b) < SCORE_TOL
    out = K.gemm_nt(ap, dev(b)).cpu().numpy()                      # mixed: the fp32 side is quantized on the way in
    assert normwise(out, ref, a, b) < SCORE_TOL


def _k_panels(K, x):
    """[rows, k] numpy -> fp16 Panels operand [ceil(k/64), rows, 64] with a zero tail."""
    return K.Panels.from_dense(dev(x))


@pytest.mark.parametrize("M,N,Kd", [(70, 40, 9000), (512, 512, 4096), (1000, 200, 64), (129, 257, 100), (33, 64, 31)])
def test_gemm_mn_major_operands(K, M, N, Kd):
    """Every operand layout of okge_gemm_f16_nt gives the same product: row-major and K-panels (K-major in shared
    memory) and the two MN-major forms the backward contractions use (E and Q read as their own transposes, the dS
    panels read as dS^T) -- including row counts that are not multiples of 64 (per-box TMA path, zero-filled edges).
    The fp32 (TF32) contraction takes the row-major and column-major forms."""
    rng = np.random.default_rng(M + 3 * N + Kd)
    a = rng.standard_normal((M, Kd)).astype(np.float32)
    b = rng.standard_normal((N, Kd)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    aT, bT = np.ascontiguousarray(a.T), np.ascontiguousarray(b.T)
    forms_a = {"row": K.quantize(dev(a)), "kpan": _k_panels(K, a), "col": K.ColMajor(K.quantize(dev(aT))),
               "mnpan": _k_panels(K, aT).T}
    forms_b = {"row": K.quantize(dev(b)), "kpan": _k_panels(K, b), "col": K.ColMajor(K.quantize(dev(bT))),
               "mnpan": _k_panels(K, bT).T}
    base = None
    for na, fa in forms_a.items():
        for nb, fb in forms_b.items():
            out = K.gemm_nt(fa, fb, splits=1).cpu().numpy()
            assert out.shape == (M, N)
            assert normwise(out, ref, a, b) < SCORE_TOL, (na, nb)
            if base is None:
                base = out
            # same rounded inputs, same fp32 accumulation: the layouts agree to accumulation-order noise
            assert np.abs(out - base).max() <= 1e-5 * np.abs(ref).max() + 1e-6, (na, nb)
    if Kd >= 4096:
        out = K.gemm_nt(forms_a["kpan"], forms_b["col"], splits=4).cpu().numpy()     # the dQ = dS E instance (split-K)
        assert normwise(out, ref, a, b) < SCORE_TOL
    for fa in (dev(a), K.ColMajor(dev(aT))):                                         # okge_gemm_tf32_nt
        for fb in (dev(b), K.ColMajor(dev(bT))):
            assert normwise(K.gemm_nt(fa, fb, splits=1).cpu().numpy(), ref, a, b) < SCORE_TOL


@pytest.mark.parametrize("M,N,Kd,n_ids", [(1000, 200, 130, 40), (300, 512, 64, 0), (129, 36, 33, 7), (4096, 64, 512, 300),
                                          # K >= 2,048: the deep-ring instantiation (16-column p / G chunks, SWIZZLE_64B)
                                          (700, 200, 2100, 20), (300, 36, 2048, 5), (20000, 512, 4096, 300)])
def test_gemm_adagrad_matches_unfused(K, M, N, Kd, n_ids):
    """okge_gemm_adagrad == okge_gemm_f16_nt -> (+ extra rows) -> okge_adagrad_dense on the same operands: the fused
    epilogue applies torch.optim.Adagrad's update (utils/optim.py:194-201) to the gradient tile while it is still in
    tensor memory. Operands as in dE = dS^T Q: A = MN-panel view of dS [B, M], B = Q [B, N] read column-major."""
    rng = np.random.default_rng(M + N + Kd)
    dS = rng.standard_normal((Kd, M)).astype(np.float32)          # [B, n_entities]
    q = rng.standard_normal((Kd, N)).astype(np.float32)           # [B, D]
    a, b = _k_panels(K, dS).T, K.ColMajor(K.quantize(dev(q)))
    p0 = rng.standard_normal((M, N)).astype(np.float32)
    scale = dev(np.array([0.37], np.float32))
    ids = rng.integers(0, M, n_ids).astype(np.int32)              # duplicates included
    if n_ids > 3:
        ids[3] = ids[0]
    rows = rng.standard_normal((n_ids, N)).astype(np.float32)
    slot_map = torch.full((M,), -1, dtype=torch.int32, device="cuda")
    p_f, G_f = dev(p0), torch.zeros(M, N, device="cuda")
    p_u, G_u = dev(p0), torch.zeros(M, N, device="cuda")
    shadow = K.quantize(p_f) if N % 8 == 0 else None               # fp16 copy of the table, refreshed by the fused step
    for step in range(2):
        # unfused: materialise g, scatter-add the extra rows, dense step
        g = K.gemm_nt(a, b, alpha_dev=scale, splits=1).contiguous()
        if n_ids:
            K.scatter_add_rows(dev(rows), dev(ids), g)
        K.adagrad_dense(p_u, g, G_u, 0.3, 1e-8, 1e-10)
        # fused
        extra = emap = None
        if n_ids:
            extra = torch.zeros(n_ids, N, device="cuda")
            K.row_slots_build(dev(ids), slot_map)
            K.row_slots_accumulate(dev(rows), dev(ids), slot_map, extra)
            emap = slot_map
        K.gemm_adagrad(a, b, p_f, G_f, 0.3, 1e-8, 1e-10, alpha_dev=scale, extra_map=emap, extra=extra, shadow=shadow)
        if n_ids:
            K.row_slots_clear(dev(ids), slot_map)
            assert bool((slot_map == -1).all())
    if shadow is not None:                                         # == quantizing the updated table with the same scale
        scale_p = 1.0 / float(shadow.inv_scale.item())
        assert torch.equal(shadow.hi[:, :N], (p_f * scale_p).to(torch.float16))
    # same tile arithmetic and accumulator; the fused epilogue uses the MUFU sqrt / reciprocal (a few ulps of the update
    # term) and adds the few extra rows in a different order
    np.testing.assert_allclose(G_f.cpu().numpy(), G_u.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(p_f.cpu().numpy(), p_u.cpu().numpy(), rtol=1e-5, atol=2e-6)
    if n_ids == 0:
        assert torch.equal(G_f, G_u)                              # the accumulator is bit-identical to the two-kernel path
        assert np.abs(p_f.cpu().numpy() - p_u.cpu().numpy()).max() <= 3e-7 * 0.3 + 2.0 ** -23 * np.abs(p0).max()


@pytest.mark.parametrize("name", ["lookup_distmult_bce", "lookup_complex_bce_smooth"])
def test_fused_entity_update_matches_unfused_step(K, name):
    """Trainer args["fused_entity_update"]: the entity table's gradient stays factored (dS panels + Q + the batch's own
    lookup rows) and Adagrad.step applies it through okge_gemm_adagrad. Post-step weights and accumulator equal the
    unfused path (same kernels, same formula); `.grad` of the table is never materialised."""
    from open_knowledge_graph_embeddings_b200.trainer import AddLossModule
    from open_knowledge_graph_embeddings_b200.optim import OptimRegime
    case, gold = MODEL_CASES[name], load_golden(name)
    _, _, loss_name, smoothing, _, _, _ = case
    results = {}
    for fused in (False, True):
        model = build_model(case, gold)
        model.fused_entity_update = fused
        inputs, labels, N = batch_from_gold(gold, "train")
        mwl = AddLossModule(model, torch.nn.BCEWithLogitsLoss(reduction="sum"), smoothing)
        oc = {"optimizer": "Adagrad", "epoch": 0, "lr": 0.3, "weight_decay": 1e-10}
        opts = OptimRegime.setup_optimizer_regime({"optimization_config": oc, "lr_scheduler_config": None}, model)
        model.train()
        for step in range(2):
            for o in opts:
                o.update(1, step)
                o.zero_grad()
            loss, _, _ = mwl(inputs, labels, False, None, 1, "right_and_left_prefix")
            (loss.sum() / int(gold["train/normalizer_loss"])).backward()
            w = model.entity_embedding.weight
            if fused:
                assert w.grad is None and w._okge_deferred is not None
            else:
                assert w.grad is not None
            for o in opts:
                o.step()
        st = opts[0].optimizer.state[model.entity_embedding.weight]
        assert st["step"] == 2
        results[fused] = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        results[fused]["__sum"] = st["sum"].cpu().numpy()
    # The first Adagrad steps are sign-like, clr * g / (|g| + eps): where the 1-vs-all row and the lookup row of an entity
    # nearly cancel, the order in which the two paths add them (dE + dx1 + dx2 vs dE + (dx1 + dx2)) shows up at the
    # 1e-4 * lr level in a handful of elements; everything else agrees to fp32 round-off.
    for k in results[False]:
        np.testing.assert_allclose(results[True][k], results[False][k], rtol=2e-5, atol=1e-4 * 0.3, err_msg=k)
        close = np.isclose(results[True][k], results[False][k], rtol=2e-5, atol=2e-6)
        assert close.mean() > 0.99, k
    # and the fused path lands on the reference's post-step weights wherever the gradient is above the TF32 noise floor
    # (first step only in the golden; here we just check the update moved the table)
    assert np.abs(results[True]["entity_embedding.weight"] - gold["init/entity_embedding.weight"]).max() > 1e-3


@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_score_bce_vs_oracle(K, smoothing):
    rng = np.random.default_rng(4)
    B, N, D = 150, 3001, 64
    q = (0.4 * rng.standard_normal((B, D))).astype(np.float32)
    E = (0.4 * rng.standard_normal((N, D))).astype(np.float32)
    ptr, idx = random_csr(rng, B, N, 6)
    y = O.smooth_labels(O.dense_labels(ptr, idx, N), smoothing)
    scores = q.astype(np.float64) @ E.astype(np.float64).T
    y_base, y_pos = (0.0, 1.0) if smoothing == 0 else ((1 - smoothing) / N, (1 + 1 / N) * (1 - smoothing))
    loss, dS = K.score_bce(dev(q), dev(E), dev(ptr), dev(idx), y_base, y_pos)
    ref_loss = O.bce_with_logits_sum(scores, y)
    assert abs(loss.item() - ref_loss) <= LOSS_RTOL * abs(ref_loss)
    ref_dS = O.bce_with_logits_grad(scores, y.astype(np.float64))
    assert dS.shape == (B, N)                                       # fp16 K-panel operand of the dQ / dE contractions
    assert np.abs(dS.dense().cpu().numpy() - ref_dS).max() < 1e-3   # |sigmoid'| <= 1/4 times the score tolerance + 2^-11
    assert torch.all(dS.data[-1, :, N % 64:] == 0), "tail of the last panel is the zero K-padding"


@pytest.mark.parametrize("count", [1, 31, 257, 1000, 1200])
def test_score_bce_device_side_column_limit(K, count):
    """okge_score_bce with n_cols_dev: a candidate operand padded to a fixed capacity gives the loss and the gradient of
    its first `count` rows; padded columns get an exactly zero gradient whatever the padding rows hold."""
    rng = np.random.default_rng(count)
    B, cap, D = 130, 1200, 64
    q = dev((0.4 * rng.standard_normal((B, D))).astype(np.float32))
    E = dev((0.4 * rng.standard_normal((cap, D))).astype(np.float32))
    E[count:] = 3.0                                  # padding rows: any values of ordinary magnitude (stale encodes)
    ptr, idx = random_csr(rng, B, count, 4)
    n_dev = torch.tensor([count], dtype=torch.int32, device="cuda")
    q16, e16 = K.quantize(q), K.quantize(E)
    loss, dS = K.score_bce(q16, e16, dev(ptr), dev(idx), n_cols_dev=n_dev)
    ref_loss, ref_dS = K.score_bce(q16, e16.row_slice(0, count), dev(ptr), dev(idx))
    assert abs(loss.item() - ref_loss.item()) <= 1e-9 * abs(ref_loss.item())
    d = dS.dense()
    assert torch.equal(d[:, :count], ref_dS.dense())
    assert torch.all(d[:, count:] == 0)


def test_score_lse_and_softmax_grad_vs_oracle(K):
    rng = np.random.default_rng(5)
    B, N, D = 140, 2500, 64
    q = (0.5 * rng.standard_normal((B, D))).astype(np.float32)
    E = (0.5 * rng.standard_normal((N, D))).astype(np.float32)
    ptr, idx = random_csr(rng, B, N, 5)
    y = O.dense_labels(ptr, idx, N)
    scores = q.astype(np.float64) @ E.astype(np.float64).T
    lse, pos = K.score_lse(dev(q), dev(E), dev(ptr), dev(idx))
    ref_lse = -O.log_softmax_rows(scores)[:, 0] + scores[:, 0]
    bound = (np.linalg.norm(q, axis=1) * np.linalg.norm(E, axis=1).max())
    assert (np.abs(lse.cpu().numpy() - ref_lse) / bound).max() < SCORE_TOL
    loss = float((np.diff(ptr) * lse.cpu().numpy().astype(np.float64)).sum() - pos.cpu().numpy().astype(np.float64).sum())
    ref_loss = O.kl_log_softmax_sum(scores, y)
    assert abs(loss - ref_loss) <= LOSS_RTOL * abs(ref_loss)
    w = torch.from_numpy(np.diff(ptr).astype(np.float32)).cuda()
    dS = K.score_softmax_grad(dev(q), dev(E), dev(ptr), dev(idx), lse, w)
    ref = O.kl_log_softmax_grad(scores, y.astype(np.float64))
    # d(w * softmax) = w * p * d(s - lse): the score tolerance scaled by the row weight and the probability
    p_max = float(np.exp(O.log_softmax_rows(scores)).max())
    assert np.abs(dS.dense().cpu().numpy() - ref).max() <= 2 * SCORE_TOL * bound.max() * np.diff(ptr).max() * p_max + 2e-3


def test_rank_count_bit_exact_vs_oracle(K, kats):
    # the reference KAT: ranks [2, 1, 3]
    args = [kats[k] for k in ("kat/ans_row", "kat/alt_ptr", "kat/alt_idx", "kat/filt_ptr", "kat/filt_idx")]
    t, g, e = K.rank_count(dev(kats["kat/scores"]), *(dev(a) for a in args))
    assert g.tolist() == [1, 1, 0] and e.tolist() == [3, 0, 7] and np.array_equal(t.cpu().numpy(), kats["kat/true"])
    # random scores with exact ties, alternatives, empty filters, N not a multiple of 4
    rng = np.random.default_rng(6)
    B, N = 37, 1003
    scores = np.round(rng.standard_normal((B, N)), 1).astype(np.float32)     # many exact ties
    fptr, fidx = random_csr(rng, B, N, 12, empty_rows=(1, 2, B - 1))           # some prefixes without filter
    ans_row = np.sort(rng.integers(0, B, 80)).astype(np.int32)
    aptr, aidx = random_csr(rng, len(ans_row), N, 3)
    t, g, e = K.rank_count(dev(scores), dev(ans_row), dev(aptr), dev(aidx), dev(fptr), dev(fidx))
    rt, rg, re = O.rank_counts(scores, ans_row, aptr, aidx, fptr, fidx)
    assert np.array_equal(t.cpu().numpy(), rt)
    assert np.array_equal(g.cpu().numpy(), rg) and np.array_equal(e.cpu().numpy(), re)


def test_fused_rank_bit_exact_vs_oracle_on_same_scores(K):
    """Fused scoring + counting + filter correction == compute_metrics applied to the scores the same
    kernel materialises (bit-exact), including exact ties from duplicated entity rows."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    rng = np.random.default_rng(7)
    B, N, Dm = 90, 4099, 64
    q = rng.standard_normal((B, Dm)).astype(np.float32)
    E = rng.standard_normal((N, Dm)).astype(np.float32)
    E[100:140] = E[200:240]                                     # identical mentions -> tied scores
    fptr, fidx = random_csr(rng, B, N, 20)
    ans_row = np.sort(rng.integers(0, B, 200)).astype(np.int32)
    aptr, aidx = random_csr(rng, len(ans_row), N, 3)
    aidx[:40] = np.arange(100, 140)[: len(aidx[:40])]           # some answers sit on tied rows
    ps = D.PrefixScores(dev(q), dev(E))
    filt = D.CSRMatrix(dev(fptr), dev(fidx), (B, N))
    ans = D.RankedAnswers(dev(ans_row), dev(aptr), dev(aidx))
    t, g, e, _ = D.rank_answers(filt, ans, ps)
    dense = ps.dense().cpu().numpy()
    rt, rg, re = O.rank_counts(dense, ans_row, aptr, aidx, fptr, fidx)
    assert np.array_equal(t.cpu().numpy(), rt)
    assert np.array_equal(g.cpu().numpy(), rg) and np.array_equal(e.cpu().numpy(), re)
    assert re.max() > 0, "test must exercise ties"
    # unfused kernel on the same matrix agrees too, and so do the meters
    t2, g2, e2 = K.rank_count(dev(dense), dev(ans_row), dev(aptr), dev(aidx), dev(fptr), dev(fidx))
    assert torch.equal(g, g2) and torch.equal(e, e2)
    m = D.metrics_from_counts(g, e)
    om = O.metrics_from_ranks(rg + re // 2, ans_row)
    for k in ("mrr", "mr", "h1", "h3", "h10", "h50"):
        assert m[k].avg == pytest.approx(om[k].avg, rel=1e-6) and m[k].count == om[k].count


def test_optimizers_vs_oracle(K):
    rng = np.random.default_rng(8)
    p0 = rng.standard_normal((130, 36)).astype(np.float32)
    p, G = dev(p0), torch.zeros(130, 36, device="cuda")
    rp, rG = p0.copy(), np.zeros_like(p0)
    for step in range(1, 4):
        g = rng.standard_normal(p0.shape).astype(np.float32)
        K.adagrad_dense(p, dev(g), G, 0.3, 1e-8, 1e-10)
        rp, rG = O.adagrad_step(rp, g, rG, 0.3, 1e-8, 1e-10)
    np.testing.assert_allclose(p.cpu().numpy(), rp, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(G.cpu().numpy(), rG, rtol=2e-6)
    p, m, v = dev(p0), torch.zeros(130, 36, device="cuda"), torch.zeros(130, 36, device="cuda")
    rp, rm, rv = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 4):
        g = rng.standard_normal(p0.shape).astype(np.float32)
        K.adam_dense(p, dev(g), m, v, 1e-2, 0.9, 0.999, 1e-8, 1e-6, step)
        rp, rm, rv = O.adam_step(rp, g, rm, rv, 1e-2, step, eps=1e-8, weight_decay=1e-6)
    np.testing.assert_allclose(p.cpu().numpy(), rp, rtol=2e-6, atol=2e-6)
    # row-wise variants touch only the listed rows and agree with the dense step there (wd = 0)
    p, G = dev(p0), torch.zeros(130, 36, device="cuda")
    rows = np.array([3, 77, 129, 0], np.int32)
    g = rng.standard_normal((4, 36)).astype(np.float32)
    K.adagrad_rows(p, G, dev(g), dev(rows), 0.3, 1e-8, 0.0)
    exp = p0.copy()
    exp[rows], _ = O.adagrad_step(p0[rows], g, np.zeros_like(g), 0.3, 1e-8, 0.0)
    np.testing.assert_allclose(p.cpu().numpy(), exp, rtol=2e-6, atol=2e-6)


def test_dropout_statistics_and_replay(K):
    x = torch.randn(1 << 20, device="cuda")
    a, b = K.dropout(x, 0.4, seed=11, offset=0), K.dropout(x, 0.4, seed=11, offset=0)
    assert torch.equal(a, b), "same (seed, offset) must reproduce the mask (used by backward)"
    keep = (a != 0).float().mean().item()
    assert abs(keep - 0.6) < 5e-3
    assert torch.allclose(a[a != 0], (x / 0.6)[a != 0])
    assert not torch.equal(a, K.dropout(x, 0.4, seed=11, offset=1 << 38))
    assert torch.equal(K.dropout(x, 0.0, seed=1), x)


@pytest.mark.parametrize("n,L,D,V", [(5, 3, 8, 11), (300, 10, 32, 50), (4096, 10, 64, 200), (1000, 10, 512, 300)])
def test_lstm_last_state_vs_oracle(K, n, L, D, V):
    """functional.LSTMLastState (tensor-core gate products + okge_lstm_cell_*) against the numpy restatement of the
    reference's LSTM encoder: encoded rows and all five gradients. TF32 products through L recurrent steps."""
    from open_knowledge_graph_embeddings_b200 import functional as Fn
    rng = np.random.default_rng(n + D)
    table = (0.3 * rng.standard_normal((V, D))).astype(np.float32)
    tok = rng.integers(1, V, size=(n, L))
    lens = rng.integers(1, L + 1, size=n)
    lens[::7] = 0                                           # rows without any token: last_state = -1 wraps to step L - 1
    tok[np.arange(L)[None, :] >= lens[:, None]] = 0
    k = 1.0 / np.sqrt(D)
    w_ih, w_hh = (rng.uniform(-k, k, (4 * D, D)).astype(np.float32) for _ in range(2))
    b_ih, b_hh = (rng.uniform(-k, k, 4 * D).astype(np.float32) for _ in range(2))
    g = rng.standard_normal((n, D)).astype(np.float32)
    ref, cache = O.lstm_last_state_encode(table, tok, np.arange(n), w_ih, w_hh, b_ih, b_hh)
    ref_grads = O.lstm_last_state_backward(g, cache)
    params = [dev(a).requires_grad_(True) for a in (table, w_ih, w_hh, b_ih, b_hh)]
    tok_tm = dev(np.ascontiguousarray(tok.T).astype(np.int32))
    last = dev(((lens - 1) % L).astype(np.int32))
    out = Fn.LSTMLastState.apply(*params, tok_tm, last)
    out.backward(dev(g))
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= 3e-3 * np.abs(ref).max()
    for name, p, r in zip(("table", "w_ih", "w_hh", "b_ih", "b_hh"), params, ref_grads):
        assert np.abs(p.grad.cpu().numpy() - r).max() <= 5e-3 * np.abs(r).max() + 1e-7, name
    with torch.no_grad():                                   # inference path (ping-pong state, nothing saved)
        out2 = Fn.LSTMLastState.apply(*params, tok_tm, last)
    assert torch.equal(out2, out.detach())


def _bn_reference(x, gamma, beta, rm, rv, bounds, momentum=0.1, eps=1e-5):
    """torch.nn.functional.batch_norm (training) on CPU in fp64, one call per non-empty row segment, in order."""
    x = x.double().cpu().requires_grad_(True)
    gamma, beta = gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True)
    rm, rv = rm.double().cpu().clone(), rv.double().cpu().clone()
    outs, calls = [], 0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi > lo:
            outs.append(torch.nn.functional.batch_norm(x[lo:hi], rm, rv, gamma, beta, True, momentum, eps))
            calls += 1
    return x, gamma, beta, torch.cat(outs), rm, rv, calls


@pytest.mark.parametrize("n,D,bounds", [(2, 4, None), (37, 64, None), (513, 200, None), (14541, 64, None), (6000, 512, None),
                                        (48, 32, (0, 24, 48)), (512, 200, (0, 3, 512)), (512, 64, (0, 0, 512)),
                                        (512, 64, (0, 512, 512)), (700, 128, (0, 100, 350, 700))])
def test_batch_norm_rows_vs_torch(K, n, D, bounds):
    """okge_bn_train_fwd / _bwd: per-segment statistics, running-statistic updates in segment order and all three
    gradients against torch's batch_norm in fp64 (one call per segment, like the reference's separate encode calls)."""
    g = torch.Generator().manual_seed(n + D)
    x = (torch.randn(n, D, generator=g) * 0.7 + torch.randn(D, generator=g) * 3.0)      # |mean| >> std in some columns
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g)
    rm, rv = torch.randn(D, generator=g) * 0.1, torch.rand(D, generator=g) + 0.5
    dy = torch.randn(n, D, generator=g)
    bnd = (0, n) if bounds is None else bounds
    rx, rg, rb, ry, rrm, rrv, calls = _bn_reference(x, gamma, beta, rm, rv, bnd)
    (ry * dy.double()).sum().backward()
    pairs = [v for lo, hi in zip(bnd[:-1], bnd[1:]) for v in (lo, hi)]              # [begin, end) per segment
    seg = None if bounds is None else torch.tensor(pairs, dtype=torch.int32, device="cuda")
    n_seg = len(bnd) - 1
    rm_d, rv_d, nbt = rm.cuda(), rv.cuda(), torch.full((), 7, dtype=torch.int64, device="cuda")
    y, mean, invstd = K.bn_train_fwd(x.cuda(), gamma.cuda(), beta.cuda(), rm_d, rv_d, nbt, 0.1, 1e-5, seg, n_seg)
    dx, dgamma, dbeta = K.bn_train_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, invstd, seg, n_seg)
    assert int(nbt) == 7 + calls
    np.testing.assert_allclose(y.cpu().numpy(), ry.detach().numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(rm_d.cpu().numpy(), rrm.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv_d.cpu().numpy(), rrv.numpy(), rtol=1e-5, atol=1e-6)
    # dx = gamma * invstd * (dy - c1 - xhat * c2) cancels almost completely for tiny segments: the error is relative to
    # the size of the terms, not of the result
    scale = max(float(rx.grad.abs().max()), float((gamma.cuda() * invstd.max(0).values).max()) * float(dy.abs().max()))
    np.testing.assert_allclose(dx.cpu().numpy(), rx.grad.numpy(), rtol=1e-4, atol=5e-6 * scale)
    np.testing.assert_allclose(dgamma.cpu().numpy(), rg.grad.numpy(), rtol=1e-4, atol=1e-4 * float(rg.grad.abs().max()))
    np.testing.assert_allclose(dbeta.cpu().numpy(), rb.grad.numpy(), rtol=1e-4, atol=1e-4 * float(rb.grad.abs().max()))
    # eval mode: the running statistics just written
    ye = K.bn_eval_fwd(x.cuda(), gamma.cuda(), beta.cuda(), rm_d, rv_d, 1e-5)
    ref = torch.nn.functional.batch_norm(x.double(), rrm, rrv, gamma.double(), beta.double(), False, 0.1, 1e-5)
    np.testing.assert_allclose(ye.cpu().numpy(), ref.numpy(), rtol=2e-5, atol=2e-5)


def test_batch_norm_split_phases_match_fused_call(K):
    """okge_bn_col_sums / okge_bn_normalize / okge_bn_normalize_bwd (the phases a row-partitioned caller runs around its
    all-reduces) on two row blocks == okge_bn_train_fwd / _bwd on the whole operand."""
    g = torch.Generator().manual_seed(3)
    n, D = 3001, 200
    x = (torch.randn(n, D, generator=g) + 2.0).cuda()
    gamma, beta = (torch.rand(D, generator=g) + 0.5).cuda(), torch.randn(D, generator=g).cuda()
    dy = torch.randn(n, D, generator=g).cuda()
    y, mean, invstd = K.bn_train_fwd(x, gamma, beta, None, None, None, 0.1, 1e-5)
    dx, dgamma, dbeta = K.bn_train_bwd(dy, x, gamma, mean, invstd)
    blocks = [slice(0, 1234), slice(1234, n)]
    stats = sum(K.bn_col_sums(x[b]) for b in blocks)                       # the "all-reduce"
    m = stats[0] / n
    var = (stats[1] / n - m * m).clamp_min(0)
    mean2, inv2 = m.float(), 1.0 / torch.sqrt(var.float() + 1e-5)
    np.testing.assert_allclose(mean2.cpu().numpy(), mean[0].cpu().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(inv2.cpu().numpy(), invstd[0].cpu().numpy(), rtol=1e-6)
    y2 = torch.cat([K.bn_normalize(x[b], mean2, inv2, gamma, beta) for b in blocks])
    np.testing.assert_allclose(y2.cpu().numpy(), y.cpu().numpy(), rtol=1e-5, atol=1e-5)
    sums = sum(K.bn_col_sums(dy[b], x[b], mean2, inv2) for b in blocks)
    np.testing.assert_allclose(sums[0].float().cpu().numpy(), dbeta.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(sums[1].float().cpu().numpy(), dgamma.cpu().numpy(), rtol=1e-5, atol=1e-4)
    coef = (sums / n).float().contiguous()
    dx2 = torch.cat([K.bn_normalize_bwd(dy[b], x[b], mean2, inv2, coef, gamma) for b in blocks])
    np.testing.assert_allclose(dx2.cpu().numpy(), dx.cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("with_step", [False, True])
def test_batch_norm_fused_dropout_equals_separate_kernels(K, with_step):
    """okge_bn_train_fwd / _bwd with drop_p > 0 == batch norm followed by okge_dropout(_step) with the same (seed, offset
    [, step]): same mask, bit-identical forward, and the backward sees mask * dy / (1 - p)."""
    g = torch.Generator().manual_seed(8)
    n, D, p, seed, off = 777, 200, 0.3, 1234567, 5 << 38
    x = torch.randn(n, D, generator=g).cuda()
    gamma, beta = (torch.rand(D, generator=g) + 0.5).cuda(), torch.randn(D, generator=g).cuda()
    dy = torch.randn(n, D, generator=g).cuda()
    seg = torch.tensor([0, 300, 300, 777], dtype=torch.int32, device="cuda")
    step = torch.tensor(3, dtype=torch.int64, device="cuda") if with_step else None
    y, mean, invstd = K.bn_train_fwd(x, gamma, beta, None, None, None, 0.1, 1e-5, seg, 2)
    y_sep = K.dropout(y, p, seed, off, step)
    y_fused, mean_f, invstd_f = K.bn_train_fwd(x, gamma, beta, None, None, None, 0.1, 1e-5, seg, 2, dropout=(p, seed, off, step))
    assert torch.equal(y_fused, y_sep) and torch.equal(mean_f, mean) and torch.equal(invstd_f, invstd)
    assert 0.25 < float((y_fused == 0).float().mean()) < 0.35
    ref = K.bn_train_bwd(K.dropout(dy, p, seed, off, step), x, gamma, mean, invstd, seg, 2)
    got = K.bn_train_bwd(dy, x, gamma, mean, invstd, seg, 2, dropout=(p, seed, off, step))
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_batch_norm_module_semantics(K):
    """functional.batch_norm_rows keeps nn.BatchNorm1d's behaviour: single-row training batches raise, eval uses the
    running statistics, autograd delivers the gradients of x, weight and bias."""
    from open_knowledge_graph_embeddings_b200 import functional as Fn
    torch.manual_seed(0)
    bn = torch.nn.BatchNorm1d(64).cuda()
    ref = torch.nn.BatchNorm1d(64).cuda()
    x = torch.randn(50, 64, device="cuda", requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    Fn.batch_norm_rows(bn, x).square().sum().backward()
    ref(xr).square().sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(bn.weight.grad.cpu().numpy(), ref.weight.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(bn.bias.grad.cpu().numpy(), ref.bias.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(bn.running_var.cpu().numpy(), ref.running_var.cpu().numpy(), rtol=1e-5)
    assert int(bn.num_batches_tracked) == 1
    with pytest.raises(ValueError):
        Fn.batch_norm_rows(bn, torch.randn(1, 64, device="cuda"))
    bn.eval(), ref.eval()
    with torch.no_grad():
        np.testing.assert_allclose(Fn.batch_norm_rows(bn, x).cpu().numpy(), ref(x).cpu().numpy(), rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------
# the drop-in model / loss / optimizer stack vs the golden vectors of the unmodified reference
# ---------------------------------------------------------------------------------------------

def build_model(case, gold, prefix="init/"):
    from open_knowledge_graph_embeddings_b200.dataset import EntityRelationDatasetMeta
    from open_knowledge_graph_embeddings_b200.model import Models
    kind, scorer, _, _, pool, bn, _ = case
    sd = params_of(gold, prefix)
    if kind == "lookup":
        meta = EntityRelationDatasetMeta(entities_size=sd["entity_embedding.weight"].shape[0],
                                         relations_size=sd["relation_embedding.weight"].shape[0])
        name = "LookupComplexRelationModel" if scorer == "complex" else "LookupDistmultRelationModel"
        model = getattr(Models, name)(entity_slot_size=sd["entity_embedding.weight"].shape[1], init_std=0.1,
                                      train_data=meta)
    else:
        ent_rows, rel_rows = sd["entity_token_ids"], sd["relation_token_ids"]
        meta = EntityRelationDatasetMeta(
            entity_id_to_tokens_map=[[int(t) for t in r if t != 0] or [0] for r in ent_rows],
            relation_id_to_tokens_map=[[int(t) for t in r if t != 0] or [0] for r in rel_rows],
            entities_size=ent_rows.shape[0], relations_size=rel_rows.shape[0],
            entity_tokens_size=sd["entity_embedding.weight"].shape[0],
            relation_tokens_size=sd["relation_embedding.weight"].shape[0], max_length=(10, 10))
        d = sd["entity_embedding.weight"].shape[1]
        if kind == "lstm":
            name = "LSTMComplexRelationModel" if scorer == "complex" else "LSTMDistmultRelationModel"
            model = getattr(Models, name)(entity_slot_size=d, relation_slot_size=d, init_std=0.1, dropout=0.0,
                                          normalize="batchnorm" if bn else "", train_data=meta)
        else:
            model = Models.UnigramPoolingComplexRelationModel(entity_slot_size=d, relation_slot_size=d, init_std=0.1,
                                                              pool=pool, normalize="batchnorm" if bn else None,
                                                              train_data=meta)
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys    # state-dict keys are the reference's
    return model.cuda()


def batch_from_gold(gold, split):
    from open_knowledge_graph_embeddings_b200.dataset import CSRMatrix
    po = (dev(gold[f"{split}/po_rel"]).view(-1, 1), dev(gold[f"{split}/po_obj"]).view(-1, 1))
    sp = (dev(gold[f"{split}/sp_subj"]).view(-1, 1), dev(gold[f"{split}/sp_rel"]).view(-1, 1))
    B = po[0].numel() + sp[0].numel()
    N = int(gold[f"{split}/normalizer_loss"]) // B
    return [po, sp], CSRMatrix(dev(gold[f"{split}/pos_ptr"]), dev(gold[f"{split}/pos_idx"]), (B, N)), N


def score_error(case, gold, prefix, split, scores, ref_scores, training):
    """max |ds| / (||q_b|| * ||e_n||) with the operand norms taken from the oracle model."""
    kind, scorer, _, _, pool, bn, _ = case
    om = O.OracleModel(kind, scorer, params_of(gold, prefix), pool=pool, batchnorm=bn)
    Q, E = om.operands(gold[f"{split}/po_rel"], gold[f"{split}/po_obj"], gold[f"{split}/sp_subj"],
                       gold[f"{split}/sp_rel"], training=training)
    return normwise(scores, ref_scores, Q, E)


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_train_step_vs_reference_golden(K, name):
    from open_knowledge_graph_embeddings_b200.trainer import AddLossModule
    from open_knowledge_graph_embeddings_b200.optim import OptimRegime
    case, gold = MODEL_CASES[name], load_golden(name)
    _, _, loss_name, smoothing, _, _, optimizer = case
    model = build_model(case, gold)
    inputs, labels, N = batch_from_gold(gold, "train")
    loss_fn = torch.nn.BCEWithLogitsLoss(reduction="sum") if loss_name == "bce" else torch.nn.KLDivLoss(reduction="sum")
    mwl = AddLossModule(model, loss_fn, smoothing, materialize_outputs=True)
    oc = ({"optimizer": "Adagrad", "epoch": 0, "lr": 0.3, "weight_decay": 1e-10} if optimizer == "adagrad"
          else {"optimizer": "Adam", "epoch": 0, "lr": 0.01})
    opts = OptimRegime.setup_optimizer_regime({"optimization_config": oc, "lr_scheduler_config": None}, model)
    model.train()
    for o in opts:
        o.update(1, 0)
        o.zero_grad()
    # the regime inherits eps = 1e-8 from its bootstrap Adam, exactly like the reference
    assert opts[0].optimizer.param_groups[0]["eps"] == float(gold["opt/eps"]) == 1e-8
    loss, hook, scores = mwl(inputs, labels, False, None, 1, "right_and_left_prefix")
    assert hook is None
    (loss.sum() / int(gold["train/normalizer_loss"])).backward()

    ref_scores = gold["train/scores"]
    assert scores.shape == ref_scores.shape
    assert score_error(case, gold, "init/", "train", scores.detach().cpu().numpy(), ref_scores, True) < SCORE_TOL
    assert abs(loss.item() - float(gold["train/loss_sum"])) <= LOSS_RTOL * abs(float(gold["train/loss_sum"]))
    # LSTM encoders: the gradient passes through up to ten recurrent TF32 products (and the batch-norm backward) per
    # direction, the other encoders through none: 5e-3 instead of 2e-3 of max |g|
    grad_tol = 2.5 * GRAD_TOL if case[0] == "lstm" else GRAD_TOL
    for k, g in params_of(gold, "grad/").items():
        mine = dict(model.named_parameters())[k].grad.cpu().numpy()
        assert np.abs(mine - g).max() <= grad_tol * np.abs(g).max() + 1e-12, k
    # batch-norm running statistics after the reference's five encode calls (candidates, po rel, po obj, sp subj, sp rel):
    # the segmented kernels apply the momentum updates in the same order
    for k, ref in params_of(gold, "step1/").items():
        if "running_" in k or "num_batches_tracked" in k:
            # (LSTM encoders: the statistics are taken over TF32-accurate outputs)
            np.testing.assert_allclose(model.state_dict()[k].cpu().numpy(), ref, rtol=2e-4,
                                       atol=5e-5 if case[0] == "lstm" else 1e-6, err_msg=k)
    for o in opts:
        o.step()
    # Optimizer formula parity is checked tightly below with the reference's own gradients; end to end
    # the first Adagrad/Adam step is sign-like (g / (|g| + eps)), so only elements whose gradient is above
    # the TF32 noise floor are comparable.
    for k, g in params_of(gold, "grad/").items():
        mine = dict(model.named_parameters())[k].detach().cpu().numpy()
        ref = gold["step1/" + k]
        solid = np.abs(g) > 20 * GRAD_TOL * np.abs(g).max()
        if solid.any():
            assert np.abs(mine - ref)[solid].max() < 2e-2 * float(gold["opt/lr"]) + 1e-6, k


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_optimizer_step_with_reference_gradients(K, name):
    """OptimRegime -> native Adagrad/Adam on the reference's own gradients reproduces its post-step weights
    and optimizer state (eps = 1e-8 inherited, dense weight decay)."""
    from open_knowledge_graph_embeddings_b200.optim import OptimRegime
    case, gold = MODEL_CASES[name], load_golden(name)
    model = build_model(case, gold)
    optimizer = case[6]
    oc = ({"optimizer": "Adagrad", "epoch": 0, "lr": 0.3, "weight_decay": 1e-10} if optimizer == "adagrad"
          else {"optimizer": "Adam", "epoch": 0, "lr": 0.01})
    opts = OptimRegime.setup_optimizer_regime({"optimization_config": oc, "lr_scheduler_config": None}, model)
    for o in opts:
        o.update(1, 0)
    named = dict(model.named_parameters())
    for k, g in params_of(gold, "grad/").items():
        named[k].grad = dev(g)
    for o in opts:
        o.step()
    for k in params_of(gold, "grad/"):
        np.testing.assert_allclose(named[k].detach().cpu().numpy(), gold["step1/" + k], rtol=3e-6, atol=3e-7, err_msg=k)
        st = opts[0].optimizer.state[named[k]]
        key = "sum" if optimizer == "adagrad" else "exp_avg_sq"
        np.testing.assert_allclose(st[key].cpu().numpy(), gold[f"optstate/{k}/{key}"], rtol=3e-6, atol=1e-12)


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_eval_vs_reference_golden(K, name):
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import AddLossModule
    case, gold = MODEL_CASES[name], load_golden(name)
    _, _, loss_name, smoothing, _, _, _ = case
    model = build_model(case, gold, prefix="step1/")
    inputs, labels, N = batch_from_gold(gold, "eval")
    loss_fn = torch.nn.BCEWithLogitsLoss(reduction="sum") if loss_name == "bce" else torch.nn.KLDivLoss(reduction="sum")
    mwl = AddLossModule(model, loss_fn, smoothing)
    model.eval()
    with torch.no_grad():
        loss, _, pred = mwl(inputs, labels, False, None, 1, "right_and_left_prefix")
    assert isinstance(pred, D.PrefixScores)
    dense = pred.dense().cpu().numpy()
    ref = gold["eval/scores"]
    assert score_error(case, gold, "step1/", "eval", dense, ref, False) < SCORE_TOL
    assert abs(loss.item() - float(gold["eval/loss_sum"])) <= LOSS_RTOL * abs(float(gold["eval/loss_sum"]))
    B = dense.shape[0]
    filt = D.CSRMatrix(dev(gold["eval/filt_ptr"]), dev(gold["eval/filt_idx"]), (B, N))
    ans = D.RankedAnswers(dev(gold["eval/ans_row"]), dev(gold["eval/alt_ptr"]), dev(gold["eval/alt_idx"]))
    # (1) the rank kernel on the REFERENCE's scores reproduces the reference's counts bit for bit
    t, g, e = K.rank_count(dev(ref), ans.ans_row, ans.alt_ptr, ans.alt_idx, filt.ptr, filt.idx)
    assert np.array_equal(t.cpu().numpy(), gold["eval/true_score"])
    assert np.array_equal(g.cpu().numpy(), gold["eval/greater"]) and np.array_equal(e.cpu().numpy(), gold["eval/equal"])
    m_ref = D.compute_metrics(filt, ans, dev(ref))
    for k in ("mrr", "mr", "h1", "h3", "h10", "h50"):
        assert m_ref[k].avg == pytest.approx(gold[f"eval/metric/{k}"][0], rel=1e-6, abs=1e-9)
        assert m_ref[k].count == int(gold[f"eval/metric/{k}"][1])
    # (2) the fused path on our own scores == the oracle on the same scores, bit-exact
    _, g2, e2, _ = D.rank_answers(filt, ans, pred)
    _, og, oe = O.rank_counts(dense, gold["eval/ans_row"], gold["eval/alt_ptr"], gold["eval/alt_idx"],
                              gold["eval/filt_ptr"], gold["eval/filt_idx"])
    assert np.array_equal(g2.cpu().numpy(), og) and np.array_equal(e2.cpu().numpy(), oe)
    # (3) reference-format inputs (dense bool mask, list of lists of IntTensor, dense predictions) work too
    dense_mask = filt.to_dense(torch.bool)
    label_ids = [[] for _ in range(B)]
    for j, b in enumerate(gold["eval/ans_row"]):
        label_ids[b].append(torch.from_numpy(gold["eval/alt_idx"][gold["eval/alt_ptr"][j]:gold["eval/alt_ptr"][j + 1]]))
    m3 = D.compute_metrics(dense_mask, label_ids, dev(ref))
    assert m3["mrr"].avg == pytest.approx(gold["eval/metric/mrr"][0], rel=1e-6)


def test_batch_shared_entities_mode_vs_oracle(K, kats):
    """use_batch_shared_entities=True (the OLPBench training mode): candidates = entities of the batch + sampled
    negatives (openkge/dataset.py:813-868), scored through precompute_batch_shared_inputs (openkge/trainer.py:80-82)."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import AddLossModule
    gold = load_golden("lookup_complex_bce")
    model = build_model(MODEL_CASES["lookup_complex_bce"], gold)
    sizes = kats["meta/sizes"]
    idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                        kats["data/train/all_splits_entities"], int(sizes[0]), 2, is_training_data=True)
    np.random.seed(5)
    batch = D.collate_shared(idx, kats["collate_shared/train/pad/sampler"], 50)
    inputs, nl, nm, labels, _, _, shared = D.input_and_labels_to_device(batch, True, "cuda")
    assert labels.shape[1] == 50 and nl == labels.shape[0] * 50
    mwl = AddLossModule(model, torch.nn.BCEWithLogitsLoss(reduction="sum"), materialize_outputs=True)
    model.train()
    loss, _, scores = mwl(inputs, labels, True, shared, 1, "right_and_left_prefix")
    (loss / nl).backward()
    om = O.OracleModel("lookup", "complex", params_of(gold, "init/"))
    flat = lambda t: t.cpu().numpy().reshape(-1)
    ref_scores, ref_loss, ref_grads = om.loss_and_grads(flat(inputs[0][0]), flat(inputs[0][1]), flat(inputs[1][0]),
                                                        flat(inputs[1][1]), labels.ptr.cpu().numpy(),
                                                        labels.idx.cpu().numpy(), candidate_ids=flat(shared))
    assert scores.shape == ref_scores.shape == (labels.shape[0], 50)
    assert np.abs(scores.detach().cpu().numpy() - ref_scores).max() < SCORE_TOL * max(1.0, np.abs(ref_scores).max())
    assert abs(loss.item() - ref_loss) <= LOSS_RTOL * abs(ref_loss)
    g = model.entity_embedding.weight.grad.cpu().numpy()
    rg = ref_grads["entity_embedding.weight"]
    assert np.abs(g - rg).max() <= GRAD_TOL * np.abs(rg).max()


TRAJECTORIES = {"traj_lookup_complex": ("lookup", "complex", "bce", 0.0, "sum", False, "adagrad"),
                "traj_unigram_bn": ("unigram", "complex", "bce", 0.0, "sum", True, "adagrad")}


@pytest.mark.parametrize("mode", ["eager", "fused", "graph"])
@pytest.mark.parametrize("name", sorted(TRAJECTORIES))
def test_training_trajectory_vs_reference(K, kats, name, mode):
    """30 optimizer steps of the UNMODIFIED reference (tests/golden/make_golden.py: run_trajectory_case) replayed on the
    same batches from the same initial weights: the loss of EVERY step within 1e-3 relative, then the filtered evaluation
    of the whole validation split with |dMRR| < 1e-3 and Hits within one answer -- through the plain Trainer step, the
    fused entity update and the CUDA-graph replay."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    gold = load_golden(name)
    model = build_model(TRAJECTORIES[name], gold)
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    va_idx = D.PrefixIndex(kats["data/valid/seen_prefixes"], kats["data/valid/seen_entities"],
                           kats["data/valid/all_splits_entities"], int(sizes[0]), 2, False)
    bs = gold["traj/rows"].shape[1]
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=bs, device="cuda", is_training_data=True)
    valid = D.OneToNMentionRelationDataset(va_idx, meta, batch_size=bs, device="cuda", is_training_data=False)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": mode != "eager"}
    trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)
    trainer.model_with_loss.train()
    for o in trainer.optimizers:
        o.update(1, 1)
    batches = [train.collate(rows) for rows in gold["traj/rows"]]
    step = None
    if mode == "graph":
        step = trainer.make_graphed_step(batches[0], max_positives=4096, preserve_state=True)
        assert step is not None
    losses = []
    for b in batches:
        if step is not None:
            losses.append(float(step(b)) / b[1])
        else:
            trainer.compute_one_batch(b, training=True, sync_loss=False)
            losses.append(float(trainer.last_loss) / b[1])
    ref = gold["traj/loss"]
    np.testing.assert_allclose(losses, ref, rtol=LOSS_RTOL)
    assert ref[-1] < 0.2 * ref[0]                                       # the run did train
    res = trainer.evaluate(valid.get_loader(shuffle=False, drop_last=False))
    n = int(gold["eval/metric/mrr"][1])
    assert res["mrr"].count == n
    # 1e-3, or -- on the 77-answer split of the token model -- one answer moving between ranks 2 and 3 (1/6 / n) after 30
    # steps whose fp16 roundings differ from the reference's fp32 ones
    assert abs(res["mrr"].avg - gold["eval/metric/mrr"][0]) < max(1e-3, 0.17 / n)
    for k in ("h1", "h3", "h10", "h50"):
        assert abs(res[k].avg - gold[f"eval/metric/{k}"][0]) <= 1.0 / n + 1e-9, k
    assert abs(res["mr"].avg - gold["eval/metric/mr"][0]) <= 0.02 * max(gold["eval/metric/mr"][0], 1.0)


# ---------------------------------------------------------------------------------------------
# larger sizes: properties that do not need the oracle
# ---------------------------------------------------------------------------------------------

def test_large_scores_loss_and_ranks_vs_fp32_torch(K):
    """FB15k-237-sized ComplEx batch (B = 512, N = 14,541, D = 200): loss within 1e-3 of fp32 torch, fused rank
    counts == counts from the materialised scores, |dMRR| < 1e-3 against fp32 scores."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    torch.manual_seed(0)
    B, N, Dm = 512, 14541, 200
    q = 0.3 * torch.randn(B, Dm, device="cuda")
    E = 0.3 * torch.randn(N, Dm, device="cuda")
    rng = np.random.default_rng(9)
    ptr, idx = random_csr(rng, B, N, 8)
    labels = D.CSRMatrix(dev(ptr), dev(idx), (B, N))
    ref_scores = (q.double() @ E.double().t())
    ref_loss = (torch.nn.functional.softplus(ref_scores) - ref_scores * labels.to_dense(torch.float64)).sum().item()
    loss, _ = K.score_bce(q, E, labels.ptr, labels.idx, want_dS=False)
    assert abs(loss.item() - ref_loss) <= LOSS_RTOL * abs(ref_loss)
    ans = D.RankedAnswers(dev(np.arange(B, dtype=np.int32)), dev(np.arange(B + 1, dtype=np.int32)),
                          dev(idx[ptr[:-1]]))
    ps = D.PrefixScores(q, E)
    _, g, e, _ = D.rank_answers(labels, ans, ps)
    _, g2, e2 = K.rank_count(ps.dense(), ans.ans_row, ans.alt_ptr, ans.alt_idx, labels.ptr, labels.idx)
    assert torch.equal(g, g2) and torch.equal(e, e2)
    _, g3, e3 = K.rank_count(ref_scores.float(), ans.ans_row, ans.alt_ptr, ans.alt_idx, labels.ptr, labels.idx)
    mrr = D.metrics_from_counts(g, e)["mrr"].avg
    mrr_ref = D.metrics_from_counts(g3, e3)["mrr"].avg
    assert abs(mrr - mrr_ref) < 1e-3
    # split precision (the default of the ranking path): the counts are those of the fp32 scorer, up to genuine near-ties
    ranks, ranks_ref = (g + e // 2).cpu().numpy(), (g3 + e3 // 2).cpu().numpy()
    assert np.abs(ranks - ranks_ref).max() <= 1 and (ranks == ranks_ref).mean() > 0.99
    # a single fp16 pass (TF32-grade) moves ranks by a few places but not the metric
    _, g1, e1, _ = D.rank_answers(labels, ans, D.PrefixScores(q, E, split=False))
    assert abs(D.metrics_from_counts(g1, e1)["mrr"].avg - mrr_ref) < 1e-3


def test_sharded_counts_and_lse_add_up(K):
    """Entity-sharded evaluation: integer rank counts of the shards add up exactly to the unsharded counts,
    loss partials add up, and per-shard log-sum-exp merges to the global one."""
    torch.manual_seed(1)
    B, N, Dm = 200, 9000, 64
    q = torch.randn(B, Dm, device="cuda")
    E = torch.randn(N, Dm, device="cuda")
    thr = torch.randn(B, device="cuda") * 3
    q16, e16 = K.quantize(q), K.quantize(E)               # one operand, sliced by rows: the shards share its scale
    g = torch.zeros(B, dtype=torch.int32, device="cuda"); e = torch.zeros_like(g)
    K.score_rank(q16, e16, thr, g, e)
    gs = torch.zeros_like(g); es = torch.zeros_like(g)
    for lo, hi in ((0, 2304), (2304, 4608), (4608, 9000)):       # shard boundaries, one not tile aligned
        K.score_rank(q16, e16.row_slice(lo, hi), thr, gs, es)
    assert torch.equal(g, gs) and torch.equal(e, es)
    empty_ptr = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    empty_idx = torch.zeros(0, dtype=torch.int32, device="cuda")
    lse, _ = K.score_lse(q16, e16, empty_ptr, empty_idx)
    parts = torch.stack([K.score_lse(q16, e16.row_slice(lo, hi), empty_ptr, empty_idx)[0] for lo, hi in ((0, 4608), (4608, 9000))])
    assert torch.allclose(torch.logsumexp(parts, dim=0), lse, rtol=0, atol=2e-5 * lse.abs().max().item())
    full, _ = K.score_bce(q16, e16, empty_ptr, empty_idx, want_dS=False)
    halves = sum(K.score_bce(q16, e16.row_slice(lo, hi), empty_ptr, empty_idx, want_dS=False)[0]
                 for lo, hi in ((0, 4608), (4608, 9000)))
    assert abs(full.item() - halves.item()) <= 1e-9 * abs(full.item())
    # shards that quantize their own rows (their own power-of-two scale) see the same fp16 mantissas: identical counts
    gq = torch.zeros_like(g); eq_ = torch.zeros_like(g)
    for lo, hi in ((0, 4608), (4608, 9000)):
        K.score_rank(q16, K.quantize(E[lo:hi]), thr, gq, eq_)
    assert int((g - gq).abs().max()) <= 1 and int((e - eq_).abs().max()) <= 1        # (a subnormal element may round apart)


def test_trainer_loop_runs_and_learns(K, kats):
    """Trainer.compute_one_batch / train_epoch / evaluate on the tiny reference-format dataset: loss goes down,
    metrics are produced with the reference's meter counts."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    va_idx = D.PrefixIndex(kats["data/valid/seen_prefixes"], kats["data/valid/seen_entities"],
                           kats["data/valid/all_splits_entities"], int(sizes[0]), 2, False)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    valid = D.OneToNMentionRelationDataset(va_idx, meta, batch_size=32, device="cuda", is_training_data=False)
    torch.manual_seed(3)
    model = Models.LookupComplexRelationModel(entity_slot_size=32, init_std=0.1, input_dropout=0.2, train_data=meta).cuda()
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0}
    loader = train.get_loader(shuffle=True, drop_last=True)
    trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid, loader)
    first = trainer.train_epoch(train.get_loader(shuffle=True, drop_last=True, seed=0))["loss"].avg
    for ep in range(1, 6):
        last = trainer.train_epoch(train.get_loader(shuffle=True, drop_last=True, seed=ep))["loss"].avg
    assert last < 0.7 * first
    res = trainer.evaluate(valid.get_loader(shuffle=False, drop_last=False))
    n_answers = len(va_idx.alt_ptr) - 1
    assert res["mrr"].count == n_answers and 0 < res["mrr"].avg <= 1 and res["h50"].avg >= res["h1"].avg
    # evaluate() is pipelined by one batch; it must equal the batch-synchronous loop of the reference (trainer.py:363-369)
    from open_knowledge_graph_embeddings_b200.metrics import MetricResult
    sync = MetricResult()
    with torch.no_grad():
        for batch in valid.get_loader(shuffle=False, drop_last=False):
            r, _ = trainer.compute_one_batch(batch, training=False)
            sync = sync + r
    for k in ("mrr", "mr", "h1", "h3", "h10", "h50", "loss"):
        assert res[k].count == sync[k].count and res[k].avg == pytest.approx(sync[k].avg, rel=1e-9), k


def test_lagged_loss_meter_matches_synchronous(K, kats):
    """sync_loss="lagged" delivers exactly the per-step losses of the synchronous meter, one step late."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True}
    seen = {}
    for mode in (True, "lagged"):
        torch.manual_seed(5)
        model = Models.LookupDistmultRelationModel(entity_slot_size=32, init_std=0.1, train_data=meta).cuda()
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        trainer.model_with_loss.train()
        vals = []
        for step, batch in enumerate(train.get_loader(shuffle=True, drop_last=True, seed=1)):
            for o in trainer.optimizers:
                o.update(trainer.epoch, trainer.training_steps)
            r, _ = trainer.compute_one_batch(batch, training=True, sync_loss=mode)
            trainer.training_steps += 1
            if r["loss"].count:
                vals.append(r["loss"].val)
        if mode == "lagged":
            vals.append(trainer.flush_loss()["loss"].val)
        seen[mode] = vals
    assert len(seen[True]) == len(seen["lagged"]) > 2
    np.testing.assert_allclose(seen["lagged"], seen[True], rtol=1e-5)   # atomics order differs run to run


@pytest.mark.parametrize("name,pool", [("unigram_complex_sum_bn_bce", "sum"), ("unigram_complex_max_bce", "max")])
def test_candidate_sharded_unigram_model_on_device(K, name, pool):
    """sharded.CandidateShardedUnigramModel with the CUDA kernels (one rank): the reference's golden training step
    (loss; post-step tables wherever the gradient is above the noise floor) and, on the post-step golden weights, its
    filtered rank counts -- bit-equal to the oracle's counts on the scores the same kernels materialise."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.sharded import CandidateShardedUnigramModel
    gold = load_golden(name)
    inputs, labels, N = batch_from_gold(gold, "train")
    params = {k: dev(v) for k, v in params_of(gold, "init/").items()}
    model = CandidateShardedUnigramModel(params, N, 0, 1, scorer="complex", pool=pool, lr=0.3, eps=1e-8, weight_decay=1e-10)
    loss = model.train_step((inputs, int(gold["train/normalizer_loss"]), 0.0, labels, None, None, None))
    assert abs(loss.item() - float(gold["train/loss_sum"])) <= LOSS_RTOL * abs(float(gold["train/loss_sum"]))
    for k, g in params_of(gold, "grad/").items():
        solid = np.abs(g) > 20 * GRAD_TOL * np.abs(g).max()
        mine, ref = model.p[k].cpu().numpy(), gold["step1/" + k]
        if solid.any():
            assert np.abs(mine - ref)[solid].max() < 2e-2 * 0.3 + 1e-6, k
    for k in ("entity_batchnorm.running_mean", "entity_batchnorm.running_var", "relation_batchnorm.running_var"):
        if k in model.p:
            np.testing.assert_allclose(model.p[k].cpu().numpy(), gold["step1/" + k], rtol=1e-4, atol=1e-5, err_msg=k)
    # evaluation on the golden post-step weights
    params = {k: dev(v) for k, v in params_of(gold, "step1/").items()}
    model = CandidateShardedUnigramModel(params, N, 0, 1, scorer="complex", pool=pool)
    inputs, _, _ = batch_from_gold(gold, "eval")
    B = inputs[0][0].numel() + inputs[1][0].numel()
    filt = D.CSRMatrix(dev(gold["eval/filt_ptr"]), dev(gold["eval/filt_idx"]), (B, N))
    ans = D.RankedAnswers(dev(gold["eval/ans_row"]), dev(gold["eval/alt_ptr"]), dev(gold["eval/alt_idx"]))
    _, greater, equal = model.eval_counts((inputs, 0, 0.0, None, ans, filt, None))
    E, lo, hi, e16 = model.candidate_block()
    Q, _ = model._queries(inputs, False)
    dense = K.score_store(K.quantize(Q.contiguous(), split=True), e16, split=True).cpu().numpy()
    ref = gold["eval/scores"]
    om = O.OracleModel("unigram", "complex", params_of(gold, "step1/"), pool=pool, batchnorm="bn" in name)
    Qo, Eo = om.operands(gold["eval/po_rel"], gold["eval/po_obj"], gold["eval/sp_subj"], gold["eval/sp_rel"], training=False)
    assert normwise(dense, ref, Qo, Eo) < SCORE_TOL
    _, og, oe = O.rank_counts(dense, gold["eval/ans_row"], gold["eval/alt_ptr"], gold["eval/alt_idx"],
                              gold["eval/filt_ptr"], gold["eval/filt_idx"])
    assert np.array_equal(greater.cpu().numpy(), og) and np.array_equal(equal.cpu().numpy(), oe)


def test_forward_modes_over_several_waves(K):
    """4 query tiles x 149 entity tiles (the persistent grid wraps around several times): fused BCE loss + dS,
    log-sum-exp and fused rank counts against the oracle / the materialised scores; the result for a row does not depend
    on how many query tiles run beside it."""
    rng = np.random.default_rng(77)
    B, N, D = 400, 38000, 64
    q = (0.4 * rng.standard_normal((B, D))).astype(np.float32)
    E = (0.4 * rng.standard_normal((N, D))).astype(np.float32)
    ptr, idx = random_csr(rng, B, N, 5)
    scores = q.astype(np.float64) @ E.astype(np.float64).T
    y = O.dense_labels(ptr, idx, N)
    loss, dS = K.score_bce(dev(q), dev(E), dev(ptr), dev(idx))
    ref_loss = O.bce_with_logits_sum(scores, y)
    assert abs(loss.item() - ref_loss) <= LOSS_RTOL * abs(ref_loss)
    assert np.abs(dS.dense().cpu().numpy() - O.bce_with_logits_grad(scores, y.astype(np.float64))).max() < 1e-3
    lse, pos = K.score_lse(dev(q), dev(E), dev(ptr), dev(idx))
    ref_lse = -O.log_softmax_rows(scores)[:, 0] + scores[:, 0]
    bound = np.linalg.norm(q, axis=1) * np.linalg.norm(E, axis=1).max()
    assert (np.abs(lse.cpu().numpy() - ref_lse) / bound).max() < SCORE_TOL
    # rank counts: fused == counts over the materialised scores, bit-exact
    for split in (False, True):
        q16, e16 = K.quantize(dev(q), split=split), K.quantize(dev(E), split=split)
        mat = K.score_store(q16, e16, split=split)
        thr = mat[:, 17].contiguous()
        gr = torch.zeros(B, dtype=torch.int32, device="cuda")
        eq = torch.zeros(B, dtype=torch.int32, device="cuda")
        K.score_rank(q16, e16, thr, gr, eq, split=split)
        assert torch.equal(gr, (thr[:, None] < mat).sum(1).int()) and torch.equal(eq, (thr[:, None] == mat).sum(1).int())
        # 3 query tiles instead of 4: same rows, same bits
        mat_odd = K.score_store(q16.row_slice(0, 300), e16, split=split)
        assert torch.equal(mat_odd, mat[:300])


def test_full_size_properties_at_c3_shape(K):
    """BASELINE.json configs[2] at its real size (N = 10^6 candidates, D = 512, B = 512), where no oracle can run:
    size-independent properties. (a) the fused loss over all candidates equals the sum over four column blocks;
    (b) fused rank counts are additive over the same blocks, exactly (integers); (c) the fused dE + Adagrad step equals
    the two-kernel path (materialised dE, dense Adagrad); (d) dQ from the MN-major operands equals dQ from explicitly
    transposed K-major copies of a slice."""
    torch.manual_seed(11)
    N, D, B = 1_000_000, 512, 512
    E = torch.randn(N, D, device="cuda") * 0.1
    q = K.fold_query(K.FOLD_DISTMULT, torch.randn(B, D, device="cuda") * 0.1, torch.randn(B, D, device="cuda") * 0.1 + 1.0)
    ptr = torch.arange(0, 2 * B + 1, 2, dtype=torch.int32, device="cuda")
    idx = torch.sort(torch.randint(0, N, (B, 2), device="cuda"), dim=1).values.reshape(-1).to(torch.int32)
    idx[1::2] = torch.maximum(idx[1::2], idx[0::2] + 1).clamp(max=N - 1)          # ascending, unique per row
    blocks = [(0, 250_112), (250_112, 500_000), (500_000, 777_777), (777_777, N)]

    def restrict(lo, hi):
        keep = (idx >= lo) & (idx < hi)
        rows = torch.arange(B, device="cuda").repeat_interleave(2)[keep]
        p2 = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
        p2[1:] = torch.bincount(rows, minlength=B).cumsum(0)
        return p2, (idx[keep] - lo).to(torch.int32)

    q16, e16 = K.quantize(q), K.quantize(E)
    loss, dS = K.score_bce(q16, e16, ptr, idx)
    parts = sum(K.score_bce(q16, e16.row_slice(lo, hi), *restrict(lo, hi), want_dS=False)[0] for lo, hi in blocks)
    # partial sums are fp32 inside a 32-column chunk and fp64 across chunks; the block bounds cut chunks differently
    assert abs(loss.item() - parts.item()) <= 2e-7 * abs(loss.item())
    assert abs(loss.item() / (B * N) - 0.6931) < 0.05                               # ~ln 2 per score at random init
    thr = torch.zeros(B, device="cuda")
    g_all = torch.zeros(B, dtype=torch.int32, device="cuda")
    e_all = torch.zeros(B, dtype=torch.int32, device="cuda")
    K.score_rank(q16, e16, thr, g_all, e_all)
    g_sum = torch.zeros_like(g_all)
    e_sum = torch.zeros_like(e_all)
    for lo, hi in blocks:
        K.score_rank(q16, e16.row_slice(lo, hi), thr, g_sum, e_sum)
    assert torch.equal(g_all, g_sum) and torch.equal(e_all, e_sum)
    assert int(g_all.min()) > 0.3 * N and int(g_all.max()) < 0.7 * N               # ~half of the scores are positive
    # (c) fused update == materialised gradient + dense Adagrad; the fused step also leaves the table's fp16 copy current
    scale = torch.tensor([1.0 / (B * N)], device="cuda")
    p_f, G_f = E.clone(), torch.zeros_like(E)
    shadow = K.quantize(p_f)
    K.gemm_adagrad(dS.T, K.ColMajor(q16), p_f, G_f, 0.3, 1e-8, 1e-10, alpha_dev=scale, shadow=shadow)
    dE = K.gemm_nt(dS.T, K.ColMajor(q16), alpha_dev=scale, splits=1).contiguous()
    p_u, G_u = E.clone(), torch.zeros_like(E)
    K.adagrad_dense(p_u, dE, G_u, 0.3, 1e-8, 1e-10)
    assert torch.equal(G_f, G_u)
    assert float((p_f - p_u).abs().max()) <= 3e-7 * 0.3 + 2.0 ** -22
    assert torch.equal(shadow.hi, (p_f * (1.0 / shadow.inv_scale)).to(torch.float16))
    # (d) MN-major operands (no transpose in memory) == explicit K-major copies, on a 4096-candidate slice
    sl = slice(123_456, 123_456 + 4096)
    _, dS_s = K.score_bce(q16, e16.row_slice(sl.start, sl.stop), *restrict(sl.start, sl.stop))
    dq_mn = K.gemm_nt(dS_s, K.ColMajor(e16.row_slice(sl.start, sl.stop)), splits=1)
    dq_k = K.gemm_nt(dS_s, K.quantize(E[sl].t().contiguous()), splits=1)
    assert float((dq_mn - dq_k).abs().max()) <= 1e-5 * float(dq_k.abs().max())
    de_mn = K.gemm_nt(dS_s.T, K.ColMajor(q16), splits=1)
    de_k = K.gemm_nt(K.Panels.from_dense(dS_s.dense().t().contiguous(), dS_s.scale), K.quantize(q.t().contiguous()), splits=1)
    assert float((de_mn - de_k).abs().max()) <= 1e-5 * float(de_k.abs().max())


def test_sampled_rows_at_c3_shape_vs_fp64(K):
    """BASELINE.json configs[2] at its real size against an fp64 contraction on the GPU, for 16 sampled query rows and ALL
    10^6 candidates: loss (rel 1e-3), loss gradient dS, and the filtered rank counts of the split-precision path against
    the counts an exact scorer would produce."""
    torch.manual_seed(12)
    N, D, B = 1_000_000, 512, 16
    E = torch.randn(N, D, device="cuda") * 0.1
    q = K.fold_query(K.FOLD_DISTMULT, torch.randn(B, D, device="cuda") * 0.3, torch.randn(B, D, device="cuda") * 0.3 + 1.0)
    ptr = torch.arange(0, 3 * B + 1, 3, dtype=torch.int32, device="cuda")
    idx = torch.sort(torch.randperm(N, device="cuda")[:3 * B].reshape(B, 3), dim=1).values.reshape(-1).to(torch.int32)
    ref = q.double() @ E.double().t()                                                # [16, 10^6] fp64
    y = torch.zeros_like(ref)
    y[torch.arange(B, device="cuda").repeat_interleave(3), idx.long()] = 1.0
    ref_loss = float((torch.nn.functional.softplus(ref) - ref * y).sum())
    loss, dS = K.score_bce(q, E, ptr, idx)
    assert abs(loss.item() - ref_loss) <= LOSS_RTOL * abs(ref_loss)
    assert float((dS.dense().double() - (torch.sigmoid(ref) - y)).abs().max()) < 1e-3
    # scores: single pass TF32-grade, split precision fp32-grade
    bound = q.double().norm(dim=1)[:, None] * E.double().norm(dim=1)[None, :]
    q2, e2 = K.quantize(q, split=True), K.quantize(E, split=True)
    assert float(((K.score_store(q2, e2).double() - ref).abs() / bound).max()) < SCORE_TOL
    split_scores = K.score_store(q2, e2, split=True)
    assert float(((split_scores.double() - ref).abs() / bound).max()) < 4e-6
    # ranking of the first positive of every row: fused split-precision counts vs counts over the fp64 scores
    thr = ref[torch.arange(B, device="cuda"), idx[0::3].long()]
    exact_g = (ref > thr[:, None]).sum(1)
    g = torch.zeros(B, dtype=torch.int32, device="cuda")
    e = torch.zeros(B, dtype=torch.int32, device="cuda")
    K.score_rank(q2, e2, split_scores[torch.arange(B, device="cuda"), idx[0::3].long()].contiguous(), g, e, split=True)
    assert int((g.long() - exact_g).abs().max()) <= 2 + int(2e-5 * N)                # near-ties within 4e-6 of the bound
    assert torch.all(e >= 1)


def _assert_same_trained_tensor(a, b, lr, tight_fraction, name):
    """Two runs of the same few Adagrad steps. The first steps are sign-like (g / (|g| + eps)), so the ordering noise of
    the float atomics in the scatter-adds is amplified in the few elements whose gradient is at the noise level: almost
    all elements must agree to round-off, practically all to 1e-3 of a step (the losses of every step, which the callers
    compare to 1e-5, are the sharp check that the two runs follow the same trajectory)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.isclose(a, b, rtol=2e-5, atol=3e-4 * lr).mean() > tight_fraction, name
    assert np.isclose(a, b, rtol=2e-5, atol=1e-3 * lr).mean() >= 0.99, name
    # a noise-level gradient may flip its sign: the first step then differs by up to 2 lr in that element
    assert np.abs(a - b).max() <= 2.1 * lr + 2e-5 * np.abs(b).max(), name


def _make_model(model_name, meta_sizes, **extra):
    """Lookup or token-pooling model over the KAT graph; token rows are synthetic (3-10 tokens per id)."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.model import Models
    n_ent, n_rel = int(meta_sizes[0]), int(meta_sizes[1])
    if model_name.startswith("Lookup"):
        meta = D.EntityRelationDatasetMeta(entities_size=n_ent, relations_size=n_rel)
        return getattr(Models, model_name)(entity_slot_size=32, init_std=0.1, train_data=meta, **extra)
    rng = np.random.default_rng(5)

    def rows(n, vocab):
        r = rng.integers(4, vocab, size=(n, 10))
        r[np.arange(10)[None, :] >= rng.integers(3, 11, size=(n, 1))] = 0
        return r
    meta = D.EntityRelationDatasetMeta(entity_id_to_tokens_map=rows(n_ent, 50), relation_id_to_tokens_map=rows(n_rel, 20),
                                       entities_size=n_ent, relations_size=n_rel, entity_tokens_size=50,
                                       relation_tokens_size=20, max_length=(10, 10))
    return getattr(Models, model_name)(entity_slot_size=32, relation_slot_size=32, init_std=0.1, train_data=meta, **extra)


@pytest.mark.parametrize("model_name,extra", [("LookupDistmultRelationModel", {}), ("LookupComplexRelationModel", {}),
                                              ("LookupComplexRelationModel", {"batch_norm": True}),
                                              ("UnigramPoolingComplexRelationModel", {"normalize": "batchnorm"}),
                                              ("UnigramPoolingDistmultRelationModel", {"pool": "max"})])
def test_graphed_train_step_matches_eager(K, kats, model_name, extra):
    """graphed.GraphedTrainStep (one CUDA-graph launch per step) leaves the same weights, optimizer state, batch-norm
    running statistics and losses as Trainer.compute_one_batch on the same batches (ComplEx: the po / sp kind of every row
    is data in the graph; batch norm: so are the row segments of its statistics)."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.graphed import GraphedTrainStep
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True}
    batches = list(train.get_loader(shuffle=True, drop_last=True, seed=3))[:6]
    assert len({b[0][0][0].numel() for b in batches}) > 1          # the po / sp split differs from batch to batch
    out = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(9)
        model = _make_model(model_name, sizes, **extra).cuda()
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        trainer.model_with_loss.train()
        for o in trainer.optimizers:
            o.update(1, 0)
        losses = []
        if mode == "graph":
            init = {k: v.clone() for k, v in model.state_dict().items()}
            step = GraphedTrainStep(trainer, rows=32, max_positives=4096, example_batch=batches[0])
            # capture ran warm-up steps on the example batch: restore the initial state (weights and accumulators)
            model.load_state_dict(init)
            for o in trainer.optimizers:
                for st in o.optimizer.state.values():
                    st["sum"].zero_()
                    st["step"] = 0
            for b in batches:
                losses.append(float(step(b)))
        else:
            for b in batches:
                trainer.compute_one_batch(b, training=True, sync_loss=False)
                losses.append(float(trainer.last_loss))
        st = trainer.optimizers[0].optimizer.state[model.entity_embedding.weight]
        out[mode] = (losses, {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}, st["sum"].cpu().numpy(), st["step"])
    np.testing.assert_allclose(out["graph"][0], out["eager"][0], rtol=1e-4)   # same trajectory up to float-atomic ordering noise
    for k in out["eager"][1]:
        _assert_same_trained_tensor(out["graph"][1][k], out["eager"][1][k], 0.3, 0.98, k)
    np.testing.assert_allclose(out["graph"][2], out["eager"][2], rtol=1e-3, atol=1e-10)
    assert out["graph"][3] == out["eager"][3] == len(batches)


@pytest.mark.parametrize("model_name,extra", [("LookupComplexRelationModel", {}),
                                              ("UnigramPoolingComplexRelationModel", {"normalize": "batchnorm"})])
def test_graphed_train_step_batch_shared_candidates(K, kats, model_name, extra):
    """Batch-shared candidate lists (openkge/dataset.py:813-860) change length per batch: the graphed step pads them to a
    fixed capacity and keeps the count on the device. Same losses, weights and running statistics as the eager step."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.graphed import GraphedTrainStep
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True,
                                           use_batch_shared_entities=True, min_size_batch_labels=24)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0}
    np.random.seed(0)
    batches = list(train.get_loader(shuffle=True, drop_last=True, seed=3))[:6]
    counts = {int(b[6].numel()) for b in batches}
    assert len(counts) > 1 and min(counts) >= 24                   # the candidate count differs from batch to batch
    out = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(9)
        model = _make_model(model_name, sizes, **extra).cuda()
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        trainer.model_with_loss.train()
        for o in trainer.optimizers:
            o.update(1, 0)
        losses = []
        if mode == "graph":
            init = {k: v.clone() for k, v in model.state_dict().items()}
            step = GraphedTrainStep(trainer, rows=32, max_positives=4096, example_batch=batches[0], max_candidates=96)
            assert step.n_cols == 96
            model.load_state_dict(init)
            for o in trainer.optimizers:
                for st in o.optimizer.state.values():
                    st["sum"].zero_()
                    st["step"] = 0
            for b in batches:
                r, _ = step.step(b, sync_loss=True)
                losses.append(r["loss"].avg)
        else:
            for b in batches:
                r, _ = trainer.compute_one_batch(b, training=True, sync_loss=True)
                losses.append(r["loss"].avg)
        out[mode] = (losses, {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()})
    np.testing.assert_allclose(out["graph"][0], out["eager"][0], rtol=1e-4)   # same trajectory up to float-atomic ordering noise
    for k in out["eager"][1]:
        _assert_same_trained_tensor(out["graph"][1][k], out["eager"][1][k], 0.3, 0.9, k)


@pytest.mark.parametrize("scorer,max_rows", [("distmult", 3072), ("complex", 3072), ("distmult", 0)])
def test_graphed_sharded_step_matches_eager(K, kats, scorer, max_rows):
    """sharded.GraphedShardedStep (the N > 1 step as one CUDA graph; here one rank, no collectives) == train_step."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.sharded import EntityShardedLookupModel, GraphedShardedStep
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    batches = [D.input_and_labels_to_device(b, True, "cuda") for b in list(train.get_loader(shuffle=True, drop_last=True, seed=3))[:5]]
    N = int(sizes[0]) - 2
    out = {}
    for mode in ("eager", "graph"):
        g = torch.Generator().manual_seed(5)
        E = (torch.randn(N, 32, generator=g) * 0.1).cuda()
        R = (torch.randn(int(sizes[1]), 32, generator=g) * 0.1).cuda()
        model = EntityShardedLookupModel(E, R, N, 0, 1, scorer=scorer, lr=0.3, eps=1e-8, weight_decay=1e-10,
                                         fused_update_max_rows=max_rows)     # 0: the unfused large-batch update path
        if mode == "graph":
            E0, R0 = model.E.clone(), model.R.clone()
            step = GraphedShardedStep(model, 32, 4096, batches[0])   # warm-up + capture steps are undone (preserve_state)
            assert torch.equal(model.E, E0) and torch.equal(model.R, R0) and model.step_count == 0
            assert float(model.G_E.abs().max()) == 0.0 and float(model.G_R.abs().max()) == 0.0
            losses = [float(step(b)) for b in batches]
            assert model.step_count == len(batches)
        else:
            losses = [float(model.train_step(b)) for b in batches]
        out[mode] = (losses, model.E.cpu().numpy(), model.R.cpu().numpy())
    np.testing.assert_allclose(out["graph"][0], out["eager"][0], rtol=1e-4)   # same trajectory up to float-atomic ordering noise
    _assert_same_trained_tensor(out["graph"][1], out["eager"][1], 0.3, 0.97, "E")
    _assert_same_trained_tensor(out["graph"][2], out["eager"][2], 0.3, 0.97, "R")


@pytest.mark.parametrize("slots", [4, 2, 1, None])
def test_eval_single_pass_loss_and_ranking_equals_two_passes(K, kats, slots, monkeypatch):
    """Trainer.compute_one_batch(training=False): the loss and the rank counts of a batch come from ONE pass over the
    candidates (okge_score_bce_rank) when no prefix row has more than 4 ranked answers; same loss and bit-identical counts
    as the two-pass route (okge_score_bce + okge_score_rank), which batches with longer answer lists keep using."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    # slots = 1: only the first answer of a row sits on its prefix row, every further one on an extra query row of the pass
    # (the tiny graph has at most 3 answers per prefix, so the 4-slot layout alone would never overflow)
    monkeypatch.setattr(D.RankedAnswers, "SLOTS", slots)
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    ev_idx = D.PrefixIndex(kats["data/valid/seen_prefixes"], kats["data/valid/seen_entities"],
                           kats["data/valid/all_splits_entities"], int(sizes[0]), 2, False)
    valid = D.OneToNMentionRelationDataset(ev_idx, meta, batch_size=48, device="cuda", is_training_data=False)
    torch.manual_seed(2)
    model = Models.LookupComplexRelationModel(entity_slot_size=32, init_std=0.3, train_data=meta).cuda()
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3}, "lr_scheduler_config": None, "grad_clip": 0}
    trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), valid, valid)
    trainer.model_with_loss.eval()
    calls = []
    from open_knowledge_graph_embeddings_b200 import _capi
    _capi.set_call_hook(lambda name, a, phase: calls.append(name) if phase == "before" else None)
    try:
        seen_fused = seen_two_pass = 0
        with torch.no_grad():
            for batch in valid.get_loader(shuffle=False, drop_last=False):
                ans = batch[4]
                results = {}
                for route in ("auto", "two_pass"):
                    b = list(batch)
                    if route == "two_pass":
                        b[4] = D.RankedAnswers(ans.ans_row.cuda(), ans.alt_ptr.cuda(), ans.alt_idx.cuda())   # overflow unknown
                    del calls[:]
                    m, _ = trainer.compute_one_batch(tuple(b), training=False)
                    results[route] = ({k: (v.avg, v.count) for k, v in m.items()}, list(calls))
                # one pass, also for rows with more answers than slots (they ride on extra query rows of the same pass)
                assert "okge_score_bce_rank" in results["auto"][1] and "okge_score_rank" not in results["auto"][1]
                assert "okge_score_bce_rank" not in results["two_pass"][1] and "okge_score_rank" in results["two_pass"][1]
                seen_fused += 1
                seen_two_pass += ans.overflow.numel() > 0
                for k, (avg, cnt) in results["two_pass"][0].items():
                    a2, c2 = results["auto"][0][k]
                    # counts are integers: exact. The one-pass loss comes from the split-precision scores of the ranking pass,
                    # the two-pass loss from a single fp16 contraction: equal to the score tolerance
                    assert c2 == cnt and a2 == pytest.approx(avg, rel=1e-5 if k == "loss" else 1e-9, abs=1e-12), k
        # the tiny graph has up to 3 answers per prefix: batches overflow 1 or 2 slots, never 4
        assert seen_fused > 0 and (slots is None or (seen_two_pass > 0) == (slots < 3))
    finally:
        _capi.set_call_hook(None)


def test_train_epoch_with_cuda_graph_option(K, kats):
    """Trainer(args["cuda_graph"]=True).train_epoch: the step is captured lazily from the first full batch WITHOUT training
    on it (state restored after the capture) and every full batch is a graph launch; same losses and weights as the eager
    epochs. A batch with another row count is not accepted by the graphed step (the loop runs it eagerly); Adam
    (step-dependent host scalars) is not captured."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.graphed import GraphCaptureUnsupported, GraphedTrainStep
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    out = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(11)
        model = Models.LookupComplexRelationModel(entity_slot_size=32, init_std=0.1, train_data=meta).cuda()
        args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
                "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True, "cuda_graph": mode == "graph"}
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        # (drop_last like the reference's training loader: its accumulation counter only steps on full batches)
        meters = [trainer.train_epoch(train.get_loader(shuffle=True, drop_last=True, seed=e)) for e in range(2)]
        out[mode] = ([m["loss"].avg for m in meters], [m["loss"].count for m in meters],
                     {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}, trainer.training_steps)
        if mode == "graph":
            assert isinstance(trainer._graphed_step, GraphedTrainStep)
            ragged = train.collate(np.arange(19))
            full = train.collate(np.arange(32))
            assert trainer._graphed_step.accepts(full) and not trainer._graphed_step.accepts(ragged)
    assert out["graph"][1] == out["eager"][1] and out["graph"][1][1] > 0 and out["graph"][3] == out["eager"][3]
    np.testing.assert_allclose(out["graph"][0], out["eager"][0], rtol=1e-4)
    for k in out["eager"][2]:
        _assert_same_trained_tensor(out["graph"][2][k], out["eager"][2][k], 0.3, 0.9, k)
    adam = {"optimization_config": {"optimizer": "Adam", "lr": 0.01}, "lr_scheduler_config": None, "grad_clip": 0}
    trainer = Trainer(adam, Models.LookupComplexRelationModel(entity_slot_size=32, init_std=0.1, train_data=meta).cuda(),
                      torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
    for o in trainer.optimizers:
        o.update(1, 0)
    batch = next(iter(train.get_loader(shuffle=False, drop_last=True)))
    with pytest.raises(GraphCaptureUnsupported):
        GraphedTrainStep(trainer, 32, 4096, batch)


def test_graphed_train_step_dropout_and_unsupported(K, kats):
    """With dropout the replayed launches take their Philox step from a device counter: two replays of the same batch from
    the same weights draw different masks (different losses), and training still learns. The N3 hook is not capturable."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.graphed import GraphCaptureUnsupported, GraphedTrainStep
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True}
    batches = list(train.get_loader(shuffle=True, drop_last=True, seed=3))
    torch.manual_seed(4)
    model = Models.LookupComplexRelationModel(entity_slot_size=32, init_std=0.1, input_dropout=0.4, train_data=meta).cuda()
    trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
    trainer.model_with_loss.train()
    for o in trainer.optimizers:
        o.update(1, 0)
    step = GraphedTrainStep(trainer, rows=32, max_positives=4096, example_batch=batches[0])
    snap = {k: v.clone() for k, v in model.state_dict().items()}
    sums = {id(st): st["sum"].clone() for o in trainer.optimizers for st in o.optimizer.state.values()}
    first = float(step(batches[1]))
    model.load_state_dict(snap)
    for o in trainer.optimizers:
        for st in o.optimizer.state.values():
            st["sum"].copy_(sums[id(st)])
    second = float(step(batches[1]))
    assert first != second and abs(first - second) < 0.2 * abs(first)      # same data and weights, fresh dropout mask
    assert int(step.dropout_step) == 5                                       # 3 warm-ups + 2 replays (capturing does not execute)
    losses = [float(step(b)) for _ in range(8) for b in batches]
    assert np.mean(losses[-len(batches):]) < 0.7 * np.mean(losses[:len(batches)])
    n3 = Models.LookupComplexRelationModel(entity_slot_size=32, init_std=0.1, l2_reg=0.1, train_data=meta).cuda()
    trainer = Trainer(args, n3, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
    with pytest.raises(GraphCaptureUnsupported):
        GraphedTrainStep(trainer, rows=32, max_positives=4096, example_batch=batches[0])


class _ThreadComm:
    """In-process stand-in for the NCCL all-reduce: the ranks of a sharded model run as THREADS of this process on one
    GPU (same stream, kernels of the two ranks interleave) and meet at every collective. Exercises the real CUDA kernels of
    the N > 1 path — partition, -1 markers, slot maps, per-shard fp16 operands — where only one GPU is available."""

    def __init__(self, world):
        import threading
        self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)
        self.local = threading.local()
        self.on = True

    def all_reduce(self, t, op=None):
        import torch.distributed as dist
        r = self.local.rank
        torch.cuda.synchronize()
        self.slots[r] = t.clone()
        self.barrier.wait()
        parts = [self.slots[i] for i in range(self.world)]
        red = torch.stack(parts).amax(0) if op == dist.ReduceOp.MAX else torch.stack(parts).sum(0)
        self.barrier.wait()
        t.copy_(red)
        return t


@pytest.mark.parametrize("scorer", ["distmult", "complex"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranks_as_threads_equal_single_rank(K, scorer, world):
    """EntityShardedLookupModel over `world` ranks (threads + in-process all-reduce, real kernels) == one rank: losses of
    two training steps to 1e-6, post-step shards equal to the single-rank rows, filtered rank counts BIT-EQUAL."""
    import threading
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    from open_knowledge_graph_embeddings_b200.sharded import EntityShardedLookupModel, shard_bounds
    spec = S.GraphSpec("mini", 3001, 23, 40000, 2000)
    tr_idx, ev_idx, meta = S.build_indexes(spec, seed=3)
    N, Dm, Bg = spec.n_entities, 64, 96
    g = torch.Generator(device="cpu").manual_seed(11)
    E = (torch.randn(N, Dm, generator=g) * 0.3).cuda()
    R = (torch.randn(meta.relations_size, Dm, generator=g) * 0.3).cuda()
    rng = np.random.default_rng(5)
    batches = [D.input_and_labels_to_device(tr_idx.collate(rng.integers(0, len(tr_idx), Bg)), True, "cuda") for _ in range(2)]
    eb = D.input_and_labels_to_device(ev_idx.collate(rng.integers(0, len(ev_idx), Bg)), False, "cuda")
    single = EntityShardedLookupModel(E.clone(), R.clone(), N, 0, 1, scorer=scorer, group="local")
    _, g1, e1 = single.eval_counts(eb)                       # ranking first: on identical weights the counts are bit-equal
    ref_losses = [float(single.train_step(b)) for b in batches]
    comm = _ThreadComm(world)
    out, errors = {}, []

    def run(rank):
        try:
            torch.cuda.set_device(0)
            comm.local.rank = rank
            lo, hi = shard_bounds(N, world, rank)
            m = EntityShardedLookupModel(E[lo:hi].clone(), R.clone(), N, rank, world, scorer=scorer, group="local")
            m.comm = comm
            _, gN, eN = m.eval_counts(eb)
            losses = [float(m.train_step(b)) for b in batches]
            out[rank] = (losses, m.E.clone(), gN.clone(), eN.clone(), lo, hi)
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)
            comm.barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for rank in range(world):
        losses, Er, gN, eN, lo, hi = out[rank]
        assert torch.equal(gN, g1) and torch.equal(eN, e1), rank
        assert losses[0] == pytest.approx(ref_losses[0], rel=1e-6) and losses[1] == pytest.approx(ref_losses[1], rel=1e-5)
        # Not bit-equal after two steps: the all-reduced dQ differs from the single-rank sum in the last bit, which can flip
        # the fp16 rounding of a few elements of the second step's query operand (one fp16 ulp = 5e-4 of a gradient term)
        d = (Er - single.E[lo:hi]).abs()
        assert float(d.max()) <= 5e-3 * 0.3 and float((d <= 1e-4 * 0.3).float().mean()) > 0.99, (rank, float(d.max()))


def test_row_wise_optimizers_with_slot_map_and_sparse_embedding(K, kats):
    """okge_adagrad_rows / okge_adam_rows with repeated ids (slot-owner convention) == the dense kernels on the scattered
    gradient (weight decay 0); a Lookup model built with sparse=True (openkge/model.py:390-391) hands its relation-table
    gradient to Adagrad as a RowsGrad and lands on the dense model's post-step weights; like torch.optim, Adagrad refuses
    sparse gradients with weight decay and Adam refuses them altogether."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import functional as Fn
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.optim import Adagrad, Adam
    rng = np.random.default_rng(21)
    p0 = rng.standard_normal((50, 24)).astype(np.float32)
    ids = np.array([3, 7, 3, 49, 0, 7, 7, 12], np.int32)
    rows = rng.standard_normal((len(ids), 24)).astype(np.float32)
    dense = np.zeros_like(p0)
    np.add.at(dense, ids, rows)
    for kind in ("adagrad", "adam"):
        p_r, p_d = dev(p0), dev(p0)
        st_r = [torch.zeros(50, 24, device="cuda") for _ in range(2)]
        st_d = [torch.zeros(50, 24, device="cuda") for _ in range(2)]
        slot_map = torch.full((50,), -1, dtype=torch.int32, device="cuda")
        summed = torch.zeros(len(ids), 24, device="cuda")
        K.row_slots_build(dev(ids), slot_map, -1)
        K.row_slots_accumulate(dev(rows), dev(ids), slot_map, summed, -1)
        if kind == "adagrad":
            K.adagrad_rows(p_r, st_r[0], summed, dev(ids), 0.3, 1e-8, 0.0, slot_map=slot_map)
            K.adagrad_dense(p_d, dev(dense), st_d[0], 0.3, 1e-8, 0.0)
        else:
            K.adam_rows(p_r, st_r[0], st_r[1], summed, dev(ids), 1e-2, 0.9, 0.999, 1e-8, 0.0, 1, slot_map=slot_map)
            K.adam_dense(p_d, dev(dense), st_d[0], st_d[1], 1e-2, 0.9, 0.999, 1e-8, 0.0, 1)
        touched = np.unique(ids)
        np.testing.assert_allclose(p_r.cpu().numpy()[touched], p_d.cpu().numpy()[touched], rtol=2e-6, atol=2e-6)
        untouched = np.setdiff1d(np.arange(50), touched)
        assert np.array_equal(p_r.cpu().numpy()[untouched], p0[untouched])           # rows without gradient do not move
    # model level
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    out = {}
    for sparse in (False, True):
        torch.manual_seed(4)
        model = Models.LookupDistmultRelationModel(entity_slot_size=32, init_std=0.1, sparse=sparse, train_data=meta).cuda()
        opt = Adagrad([model.relation_embedding.weight], lr=0.3, eps=1e-8, weight_decay=0.0)
        rel = torch.tensor([2, 5, 2, 3], dtype=torch.int32, device="cuda")
        x = model.encode_rel(rel.view(-1, 1))
        (x * torch.arange(1, 5, device="cuda").view(-1, 1)).sum().backward()
        w = model.relation_embedding.weight
        assert (w.grad is None and isinstance(w._okge_deferred, Fn.RowsGrad)) if sparse else w.grad is not None
        opt.step()
        out[sparse] = w.detach().cpu().numpy().copy()
    np.testing.assert_allclose(out[True], out[False], rtol=2e-6, atol=2e-6)
    w = model.relation_embedding.weight
    model.encode_rel(torch.tensor([[2]], dtype=torch.int32, device="cuda")).sum().backward()
    with pytest.raises(RuntimeError, match="not compatible with sparse gradients"):
        Adagrad([w], lr=0.3, weight_decay=1e-10).step()
    model.encode_rel(torch.tensor([[2]], dtype=torch.int32, device="cuda")).sum().backward()
    with pytest.raises(RuntimeError, match="does not support sparse gradients"):
        Adam([w], lr=0.01).step()


@pytest.mark.parametrize("model_name,extra", [("LookupComplexRelationModel", {}),
                                              ("UnigramPoolingComplexRelationModel", {"normalize": "batchnorm"})])
def test_device_collate_graph_matches_host_collate(K, kats, model_name, extra):
    """Batch-shared training with the collate ON THE DEVICE inside the step's CUDA graph (dataset.DeviceSharedCollate via
    Trainer.train_epoch over a row loader) == the host collate_shared feeding the same graphed step: without negative
    sampling the candidate SETS are identical (only their order differs), so the losses agree to summation order; with
    negatives the candidate count equals min_size_batch_labels and training proceeds."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True, "cuda_graph": True}
    rng = np.random.default_rng(3)
    row_sets = [rng.permutation(len(tr_idx))[:32] for _ in range(4)]
    out = {}
    for mode in ("host", "device"):
        torch.manual_seed(9)
        train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True,
                                               use_batch_shared_entities=True, min_size_batch_labels=0)
        model = _make_model(model_name, sizes, **extra).cuda()
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        trainer.model_with_loss.train()
        for o in trainer.optimizers:
            o.update(1, 1)
        host_batches = [train.collate(r) for r in row_sets]
        losses = []
        if mode == "host":
            g = trainer._graphed_step_for(host_batches[0])
            for b in host_batches:
                r, _ = g.step(b, sync_loss=True)
                losses.append(r["loss"].val)                              # loss / (B * candidates) of the step
        else:
            for r_ in row_sets:
                b = D.DeviceRows(torch.from_numpy(r_).cuda())
                r, _ = trainer._graphed_step_for_rows(b).step_rows(b, sync_loss=True)
                losses.append(r["loss"].val)
            g = trainer._graphed_step
            assert int(g.collate.overflow) == 0 and int(g.collate.nnz_total) == sum(int(b[3].idx.numel()) for b in host_batches)
        out[mode] = (losses, {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()})
    np.testing.assert_allclose(out["device"][0], out["host"][0], rtol=1e-4)
    for k, v in out["host"][1].items():
        if v.dtype.kind == "f":
            _assert_same_trained_tensor(out["device"][1][k], v, 0.3, 0.9, k)
    # with negatives: every batch has exactly min_size candidates, the run trains
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True,
                                           use_batch_shared_entities=True, min_size_batch_labels=50)
    torch.manual_seed(9)
    model = _make_model(model_name, sizes, **extra).cuda()
    trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
    first = trainer.train_epoch(train.get_row_loader(shuffle=True, seed=1))["loss"].avg
    for ep in range(4):
        last = trainer.train_epoch(train.get_row_loader(shuffle=True, seed=2 + ep))["loss"].avg
    g = trainer._graphed_step
    assert int(g.collate.overflow) == 0 and int(g.cand_count) == 50
    assert last < 0.8 * first


def _random_train_index(rng, n_prefix, n_entities, max_len):
    from open_knowledge_graph_embeddings_b200 import dataset as D
    lens = rng.integers(1, max_len + 1, n_prefix)
    lens[rng.integers(0, n_prefix, max(n_prefix // 50, 1))] = max_len * 8              # a few very popular prefixes
    lens = np.minimum(lens, n_entities)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.sort(rng.choice(n_entities, int(n), replace=False)) for n in lens]).astype(np.int32)
    prefix = rng.integers(2, 1000, (n_prefix, 2)).astype(np.int32)
    slot = (2 * rng.integers(0, 2, n_prefix)).astype(np.int32)
    return D.PrefixIndex.from_csr(prefix, slot, ptr, idx, n_cols=n_entities, offset=2, is_training_data=True)


@pytest.mark.parametrize("B,n_entities,min_size", [(64, 1000, 0), (1500, 70001, 4096), (4096, 40000, 4096), (7, 33, 64)])
def test_collate_shared_kernel_vs_host_collate(K, B, n_entities, min_size):
    """okge_collate_shared == the host restatement of the reference collate (dataset.collate_shared, checked against the
    reference's goldens in tests/test_oracle_golden.py) as SETS: same rows in the same po-first order, the same (row,
    entity) label pairs, the positives of the batch as the head of the candidate list; the negatives are distinct, not
    positives, inside the id range, change from call to call and fill the list to min_size. Integer work: exact."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    rng = np.random.default_rng(B)
    index = _random_train_index(rng, 3000, n_entities, 6)
    rows = rng.integers(0, len(index), B)                                  # with repeats
    host = D.collate_shared(index, rows, 0)
    (po, sp), labels_h, shared_h = host[0], host[3], host[6]
    n_u = int(shared_h.numel())
    coll = D.DeviceSharedCollate(index, min_size, cap_nnz=int(labels_h.idx.numel()) + 100, cap_cols=max(n_u, min_size) + 50,
                                 device="cuda")
    out = {k: v.clone() for k, v in coll(torch.from_numpy(rows).cuda()).items()}
    torch.cuda.synchronize()
    b_po = int(out["b_po"])
    assert b_po == (0 if po is None else po[0].numel())
    ent_h = torch.cat([t for t in ([po[1]] if po is not None else []) + ([sp[0]] if sp is not None else [])])
    rel_h = torch.cat([t for t in ([po[0]] if po is not None else []) + ([sp[1]] if sp is not None else [])])
    assert torch.equal(out["ent"].cpu(), ent_h.int()) and torch.equal(out["rel"].cpu(), rel_h.int())
    assert out["is_po"].cpu().tolist() == [1] * b_po + [0] * (B - b_po)
    assert int(out["nnz"]) == labels_h.idx.numel() and torch.equal(out["ptr"].cpu(), labels_h.ptr)
    cand = out["cand"].view(-1).cpu().numpy().astype(np.int64)
    want = min(max(n_u, min_size), n_entities)
    count = int(out["count"])
    assert count == want and int(coll.overflow) == 0 and int(coll.nnz_total) == labels_h.idx.numel()
    assert abs(float(out["inv_norm"]) * B * count - 1.0) < 1e-6
    pos = cand[:n_u]
    assert np.array_equal(pos, np.sort(shared_h.view(-1).numpy().astype(np.int64)))        # ascending by entity id
    neg = cand[n_u:count]
    assert len(np.unique(neg)) == len(neg) and not np.isin(neg, pos).any()
    assert neg.size == 0 or (neg.min() >= 2 and neg.max() < n_entities + 2)
    assert (cand[count:] == 2).all()
    # labels: the same (row, entity) pairs, ascending columns within every row
    idx_d, ptr_d = out["idx"].cpu().numpy(), out["ptr"].cpu().numpy()
    nnz = int(out["nnz"])
    assert (idx_d[nnz:] == -1).all()
    row_of = np.repeat(np.arange(B), np.diff(ptr_d))
    ent_d = cand[idx_d[:nnz]]
    ent_hh = shared_h.view(-1).numpy().astype(np.int64)[labels_h.idx.numpy()]
    key_d = np.sort(row_of * (n_entities + 2) + ent_d)
    key_h = np.sort(np.repeat(np.arange(B), np.diff(labels_h.ptr.numpy())) * (n_entities + 2) + ent_hh)
    assert np.array_equal(key_d, key_h)
    inside = np.ones(nnz, bool)
    inside[ptr_d[:-1][ptr_d[:-1] < nnz]] = False
    assert not (inside[1:] & (idx_d[1:nnz] <= idx_d[:nnz - 1])).any()
    # a second call draws other negatives (the Philox key is the call counter) and leaves no state behind
    out2 = coll(torch.from_numpy(rows).cuda())
    cand2 = out2["cand"].view(-1).cpu().numpy().astype(np.int64)
    assert np.array_equal(cand2[:n_u], pos) and int(out2["count"]) == want
    if count - n_u > 8 and n_entities > 4 * count:
        assert not np.array_equal(cand2[n_u:count], neg)
    assert int((coll.ws["first_draw"] != 2 ** 31 - 1).sum()) == 0


def test_collate_shared_kernel_capacities_cut_and_count(K):
    """Batches that do not fit the fixed capacities are cut (memory safe) and counted: labels beyond cap_nnz, candidates
    beyond cap_cols; the label columns that were cut point behind the list (they match no column)."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    rng = np.random.default_rng(5)
    index = _random_train_index(rng, 500, 5000, 8)
    rows = rng.integers(0, len(index), 256)
    host = D.collate_shared(index, rows, 0)
    total, n_u = int(host[3].idx.numel()), int(host[6].numel())
    coll = D.DeviceSharedCollate(index, 0, cap_nnz=total, cap_cols=n_u // 2, device="cuda")
    out = coll(torch.from_numpy(rows).cuda())
    assert int(coll.overflow) == 1 and int(out["count"]) == n_u // 2 and int(coll.scalars[K.COLLATE_N_UNIQUE]) == n_u
    pos = np.sort(host[6].view(-1).numpy())
    assert np.array_equal(out["cand"].view(-1).cpu().numpy(), pos[:n_u // 2])
    idx = out["idx"].cpu().numpy()
    assert idx.max() == n_u - 1 and (idx >= 0).all()
    coll = D.DeviceSharedCollate(index, 0, cap_nnz=total // 2, cap_cols=n_u, device="cuda")
    out = coll(torch.from_numpy(rows).cuda())
    assert int(coll.overflow) == 1 and int(out["nnz"]) == total // 2 and int(out["ptr"][-1]) == total // 2
    assert int(out["count"]) <= n_u and int(coll.scalars[7]) == total
    with pytest.raises(ValueError):                      # answer lists must be ascending
        bad = D.PrefixIndex.from_csr(index.prefix[:1], index.slot[:1], np.array([0, 2]), np.array([5, 3], np.int32), n_cols=10,
                                     offset=2, is_training_data=True)
        D.DeviceSharedCollate(bad, 0, 8, 8, "cuda")


# ---------------------------------------------------------------------------------------------
# the reference's REAL FB15k-237 fixture: filtered ranks of a whole split against the unmodified reference (fp32, CPU)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case,model_name,cfg", [
    ("lookup_complex", "LookupComplexRelationModel", dict(entity_slot_size=64, init_std=0.1)),
    ("unigram_complex", "UnigramPoolingComplexRelationModel",
     dict(entity_slot_size=64, relation_slot_size=64, init_std=0.1, pool="sum", dropout=0.0))])
def test_real_fb15k237_fixture_ranks_vs_reference(K, tmp_path, case, model_name, cfg):
    """The reference's own FB15k-237 id files (valid split of tests/golden/real_fixture.py: 10,000 triples, 19,998 ranked
    answers against all 14,541 entities, real token maps), weights drawn on both sides from the same seeded numpy generator:
    the (greater, equal) counts of every ranked answer against the reference's fp32 CPU evaluation (make_golden.py:
    run_real_fixture_case). The split-precision fp16 scorer (error ~4e-7 of |q||e|) may swap an answer with a candidate
    whose fp32 score is within a few ulps: >= 99 % of the ranks identical, none off by more than 3, |dMRR| < 1e-5 (1e-3 is
    the bar), Hits@k identical up to two answers. Single-pass fp16 (without the lo planes) stays within |dMRR| < 1e-3."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import real_fixture
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import dataset_build as B
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    gold = np.load(os.path.join(os.path.dirname(real_fixture.__file__), "real_fb15k237.npz"))
    root = real_fixture.stage(str(tmp_path / "fb"))
    meta = B.load_meta(root)
    va_idx = B.load_prefix_index(root, "valid.txt", is_training_data=False, meta=meta, exact_set_order=True)
    tr_idx = B.load_prefix_index(root, "train.txt", is_training_data=True, meta=meta)
    model = getattr(Models, model_name)(train_data=meta, **cfg)
    floats = [(k, tuple(v.shape)) for k, v in model.state_dict().items() if v.dtype.is_floating_point]
    assert sorted(k for k, _ in floats) == gold[f"{case}/keys"].tolist()                 # same parameters as the reference's model
    weights = real_fixture.seeded_weights(floats)
    model.load_state_dict({k: torch.from_numpy(w) for k, w in weights.items()}, strict=False)
    model = model.cuda()
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=256, device="cuda", is_training_data=True)
    valid = D.OneToNMentionRelationDataset(va_idx, meta, batch_size=256, device="cuda", is_training_data=False)
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.1, "weight_decay": 0.0}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0}
    trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)
    g_ref, e_ref = gold[f"{case}/greater"].astype(np.int64), gold[f"{case}/equal"].astype(np.int64)
    n = len(g_ref)

    def ranks(split):
        model.eval_split_precision = split
        model.eval()
        if hasattr(model, "precompute_embeddings_from_tokens") and not model_name.startswith("Lookup"):
            with torch.no_grad():
                model.precompute_embeddings_from_tokens()
        greater, equal = [], []
        mwl = trainer.model_with_loss
        with torch.no_grad():
            for b in valid.get_loader(shuffle=False, drop_last=False):
                inputs, nl, nm, labels, label_ids, filt, shared = valid.input_and_labels_to_device(b, training=False, device="cuda")
                mwl.defer_eval_loss = True
                try:
                    _, _, pred = mwl(inputs=inputs, labels=labels, batch_shared_entities=shared, use_batch_shared_entities=False,
                                     epoch=1, input_style_triple_or_prefix=valid.input_style)
                finally:
                    mwl.defer_eval_loss = False
                _, g, e, _ = D.rank_answers(filt, label_ids, pred)
                greater.append(g.cpu().numpy().astype(np.int64)), equal.append(e.cpu().numpy().astype(np.int64))
        return np.concatenate(greater), np.concatenate(equal)

    def mrr(g, e):
        return float(D.metrics_from_counts(torch.from_numpy(g), torch.from_numpy(e))["mrr"].avg)

    g, e = ranks(True)
    assert len(g) == n == int(gold[f"{case}/metric/mrr"][1])
    same = (g == g_ref) & (e == e_ref)
    assert same.mean() >= 0.99, same.mean()
    assert np.abs(g - g_ref).max() <= 3 and np.abs(e - e_ref).max() <= 3
    assert abs(mrr(g, e) - gold[f"{case}/metric/mrr"][0]) < 1e-5
    res = D.metrics_from_counts(torch.from_numpy(g), torch.from_numpy(e))
    for k in ("h1", "h3", "h10", "h50"):
        assert abs(res[k].avg - gold[f"{case}/metric/{k}"][0]) <= 2.0 / n + 1e-12, k
    assert abs(res["mr"].avg - gold[f"{case}/metric/mr"][0]) <= 1e-3 * gold[f"{case}/metric/mr"][0]
    g1, e1 = ranks(False)                                     # single fp16 pass: the fast evaluation mode
    assert abs(mrr(g1, e1) - gold[f"{case}/metric/mrr"][0]) < 1e-3
    assert (np.abs(g1 - g_ref) <= 0.02 * 14541).all()
    # the public entry point reports the same meters
    model.eval_split_precision = True
    total = trainer.evaluate(valid.get_loader(shuffle=False, drop_last=False))
    assert total["mrr"].count == n and abs(total["mrr"].avg - gold[f"{case}/metric/mrr"][0]) < 1e-5


# ---------------------------------------------------------------------------------------------
# N > 1 behind the public API: Trainer in a torch.distributed job (two processes on this GPU, gloo all-reduces)
# ---------------------------------------------------------------------------------------------

def _trainer_rank(rank, world, port, path, model_name, cfg, opt, loss_name, clip, graph):
    """One rank of a torch.distributed job that uses nothing but the public API: Models.<name> + Trainer (which shards the
    entity table when world > 1). Real CUDA kernels; the ranks share GPU 0 and meet in gloo all-reduces on CUDA tensors."""
    import torch.distributed as dist
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    from open_knowledge_graph_embeddings_b200.model import Models
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = S.GraphSpec("mini", 3001, 23, 40000, 2000)
    tr_idx, ev_idx, meta = S.build_indexes(spec, seed=3)
    Bg = 96
    torch.manual_seed(11)
    model = getattr(Models, model_name)(entity_slot_size=64, init_std=0.3, train_data=meta, **cfg).cuda()
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=Bg, device="cuda", is_training_data=True)
    valid = D.OneToNMentionRelationDataset(ev_idx, meta, batch_size=Bg, device="cuda", is_training_data=False)
    args = {"optimization_config": dict(opt), "lr_scheduler_config": None, "bce_label_smoothing": 0.0, "grad_clip": clip,
            "fused_entity_update": True}
    loss = torch.nn.BCEWithLogitsLoss(reduction="sum") if loss_name == "bce" else torch.nn.KLDivLoss(reduction="sum")
    trainer = Trainer(args, model, loss, train, valid)
    assert (model._shard is not None) == (world > 1)
    for o in trainer.optimizers:
        o.update(1, 1)
    rng = np.random.default_rng(5)
    batches = [D.input_and_labels_to_device(tr_idx.collate(rng.integers(0, len(tr_idx), Bg)), True, "cuda") for _ in range(3)]
    ev_rows = rng.integers(0, len(ev_idx), Bg)
    res0 = trainer.evaluate([ev_idx.collate(ev_rows)])                      # ranking first: identical weights on every layout
    inputs, nl, nm, labels, label_ids, filt, shared = D.input_and_labels_to_device(ev_idx.collate(ev_rows), False, "cuda")
    model.eval()
    with torch.no_grad():
        _, _, pred = trainer.model_with_loss(inputs=inputs, labels=labels, batch_shared_entities=shared,
                                             use_batch_shared_entities=False, epoch=1,
                                             input_style_triple_or_prefix=valid.input_style)
        _, greater, equal, _ = D.rank_answers(filt, label_ids, pred)
    trainer.model_with_loss.train()
    losses = []
    step = trainer.make_graphed_step(batches[0], max_positives=4096, preserve_state=True) if graph else None
    assert (step is not None) == bool(graph and world == 1)     # gloo collectives cannot be captured (NCCL ones are: bench)
    for b in batches:
        if step is not None:
            losses.append(float(step(b)))
        else:
            trainer.compute_one_batch(b, training=True, sync_loss=False)
            losses.append(float(trainer.last_loss))
    out = {"losses": losses, "greater": greater.cpu(), "equal": equal.cpu(), "mrr": res0["mrr"].avg, "eval_loss": res0["loss"].avg,
           "E": model.gather_entity_table().cpu(), "R": model.relation_embedding.weight.data.cpu(),
           "block": tuple(model.entity_embedding.weight.shape)}
    torch.save(out, f"{path}.{world}.{rank}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("model_name,cfg,opt,loss_name,clip,graph", [
    ("LookupDistmultRelationModel", {}, {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "bce", 0, False),
    ("LookupComplexRelationModel", {"input_dropout": 0.4}, {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "bce", 0, True),
    ("LookupComplexRelationModel", {}, {"optimizer": "Adam", "lr": 0.01}, "kl", 1.0, False),
])
def test_trainer_shards_the_entity_table_in_a_distributed_job(K, tmp_path, model_name, cfg, opt, loss_name, clip, graph):
    """The N > 1 path behind the reference's API: the same ``Models`` class and ``Trainer``, run as a 2-rank
    torch.distributed job (two processes on this GPU, gloo), against the 1-rank run from the same seed. The Trainer
    partitions the entity table; dropout (C1's input_dropout 0.4), KL + Adam + gradient clipping and the CUDA-graph replay
    all run sharded. Rank counts BIT-EQUAL, losses of three steps to 1e-5, gathered table and replicated relation table
    equal to the 1-rank run within the reduced-precision tolerance of a step."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    path = str(tmp_path / "res")
    ctx_args = (path, model_name, cfg, opt, loss_name, clip, graph)
    mp.spawn(_trainer_rank, args=(1, port, *ctx_args), nprocs=1, join=True)
    mp.spawn(_trainer_rank, args=(2, port + 1 if port < 65000 else port - 1, *ctx_args), nprocs=2, join=True)
    one = torch.load(f"{path}.1.0")
    two = [torch.load(f"{path}.2.{r}") for r in range(2)]
    lr = opt["lr"]
    dropout = cfg.get("input_dropout", 0) > 0
    assert two[0]["block"][0] + two[1]["block"][0] == one["block"][0] + 2          # PAD / UNK rows on both ranks
    for r in range(2):
        assert torch.equal(two[r]["greater"], one["greater"]) and torch.equal(two[r]["equal"], one["equal"]), r
        assert two[r]["mrr"] == one["mrr"] and two[r]["eval_loss"] == pytest.approx(one["eval_loss"], rel=1e-5)
        if dropout:
            # the candidate rows draw their dropout mask per block: another (equally valid) draw than the 1-rank run, so the
            # losses agree statistically only; both ranks of the job must still agree with each other exactly
            assert two[r]["losses"] == pytest.approx(one["losses"], rel=0.05)
            assert two[r]["losses"] == two[0]["losses"]
            continue
        assert two[r]["losses"] == pytest.approx(one["losses"], rel=1e-5), r
        d = (two[r]["E"] - one["E"]).abs()
        if opt["optimizer"] == "Adam":
            # Adam's first steps move every element by ~lr * sign(g): where |g| is of the order of eps = 1e-8 the rounding
            # noise of the all-reduce order decides the sign, so only a quantile is meaningful
            assert float((d <= 1e-2 * lr).float().mean()) > 0.99, (r, float((d <= 1e-2 * lr).float().mean()))
            continue
        assert float(d.max()) <= 5e-3 * lr and float((d <= 1e-4 * lr).float().mean()) > 0.99, (r, float(d.max()))
        assert float((two[r]["R"] - one["R"]).abs().max()) <= 5e-3 * lr
    # the replicated relation tables move in lock-step up to the summation order of the scatter-add atomics (the same
    # run-to-run noise a single GPU has); Trainer.train_epoch re-synchronises the replicas periodically
    assert torch.equal(two[0]["E"], two[1]["E"])
    # (a few fp32 ulps of the largest element: Adam's 0.01 steps leave less than that as head-room under 1e-5 lr)
    ulps = 8 * float(torch.finfo(torch.float32).eps) * float(two[0]["R"].abs().max())
    assert float((two[0]["R"] - two[1]["R"]).abs().max()) <= max(1e-5 * lr, ulps)



# ---------------------------------------------------------------------------------------------
# Data-parallel Trainer step for batch-shared candidate lists (every rank its own batch + candidate list)
# ---------------------------------------------------------------------------------------------

def _dp_setup(kats, model_name, extra, batch_size_for_backward=None):
    from open_knowledge_graph_embeddings_b200 import dataset as D
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True,
                                           use_batch_shared_entities=True, min_size_batch_labels=24)
    if batch_size_for_backward is not None:
        train.batch_size_for_backward = batch_size_for_backward
    torch.manual_seed(9)
    model = _make_model(model_name, sizes, **extra).cuda()
    np.random.seed(0)
    batches = list(train.get_loader(shuffle=True, drop_last=True, seed=3))[:6]
    return train, model, batches


_DP_ARGS = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0}


def _dp_rank(rank, world, port, path, model_name, extra, sparse=False):
    import torch.distributed as dist
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kats = load_golden("kats")
    train, model, batches = _dp_setup(kats, model_name, extra)
    trainer = Trainer(dict(_DP_ARGS, fused_entity_update=sparse), model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
    assert trainer.data_parallel and trainer.sparse_exchange == sparse
    trainer.model_with_loss.train()
    for o in trainer.optimizers:
        o.update(1, 0)
    assert trainer.make_graphed_step(batches[0], max_candidates=96) is None      # gloo all-reduces cannot be captured
    losses = []
    for b in batches[rank::world]:                       # step t: rank r trains on batch world * t + r
        r, _ = trainer.compute_one_batch(b, training=True, sync_loss=True)
        losses.append(r["loss"].avg)
    trainer.sync_replicas()                              # broadcasts rank 0's copy (incl. batch-norm running statistics)
    torch.save({"losses": losses, "state": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, f"{path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("model_name,extra,sparse", [("LookupComplexRelationModel", {}, False),
                                                     ("UnigramPoolingComplexRelationModel", {"normalize": "batchnorm"}, False),
                                                     ("UnigramPoolingComplexRelationModel", {"normalize": "batchnorm"}, True)])
def test_data_parallel_trainer_step_equals_accumulated_mean_gradient(K, kats, tmp_path, model_name, extra, sparse):
    """Batch-shared candidate lists in a 2-rank job (two processes on this GPU, gloo): every rank runs its own batch and
    candidate list, the gradients are averaged before the optimizer step (Trainer.data_parallel) -- as dense all-reduces,
    or (``sparse``: token models under ``fused_entity_update``) as the touched-row exchange: union numbering of the token
    rows either rank touches, backward into the numbered rows, one all-reduce of those rows, dense Adagrad step reading
    them through the slot map. Reference: ONE process that accumulates the gradients of the same two batches and halves
    them before the same Adagrad step. Every step's losses and the trained weights agree; both ranks end with identical
    weights."""
    import socket
    import torch.multiprocessing as mp
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    path = str(tmp_path / "dp")
    mp.spawn(_dp_rank, args=(2, port, path, model_name, extra, sparse), nprocs=2, join=True)
    two = [torch.load(f"{path}.{r}") for r in range(2)]

    train, model, batches = _dp_setup(kats, model_name, extra, batch_size_for_backward=64)
    trainer = Trainer(dict(_DP_ARGS), model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
    assert not trainer.data_parallel
    for emb in (model.entity_embedding, model.relation_embedding):
        emb.weight._okge_slot_update = False              # dense gradients, as in the data-parallel job
    trainer.data_parallel = True                          # the hook between backward and optimizer step ...
    trainer._average_gradients = lambda: [p.grad.mul_(0.5) for p in model.parameters() if p.grad is not None]   # ... halves
    trainer.model_with_loss.train()
    for o in trainer.optimizers:
        o.update(1, 0)
    ref_losses = []
    for b in batches:
        trainer.compute_one_batch(b, training=True, sync_loss=False)
        ref_losses.append(float(trainer.last_loss) / b[1])
    for r in range(2):
        np.testing.assert_allclose(two[r]["losses"], ref_losses[r::2], rtol=2e-4)
    ref_state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for k, v in ref_state.items():
        assert torch.equal(two[0]["state"][k], two[1]["state"][k]), k
        if "running_" in k or "num_batches" in k:
            continue                                      # batch-norm running statistics follow each rank's own batches
        _assert_same_trained_tensor(two[0]["state"][k].numpy(), v.numpy(), 0.3, 0.9, k)


def test_integration_md_ctypes_stub_runs(K):
    """The reference-side binding printed in INTEGRATION.md section 1 is real code: executed as written (library path
    resolved to the in-tree build) it reproduces ComplexRelationScorer._score(prefix=True, sp=True)
    (openkge/model.py:206-209, the four-mm form) in fp32-grade precision."""
    import re
    from open_knowledge_graph_embeddings_b200 import _capi
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    code = re.search(r"```python\n# openkge/b200.py.*?\n(.*?)```", text, re.S).group(1)
    code = code.replace('ctypes.CDLL("libokge_b200.so")', f'ctypes.CDLL("{_capi.LIB_PATH}")')
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    torch.manual_seed(0)
    b, d, n = 37, 64, 1001
    subj, rel, objs = (torch.randn(b, d, device="cuda"), torch.randn(b, d, device="cuda"), torch.randn(n, d, device="cuda"))
    out = ns["complex_sp_prefix_scores"](subj, rel, objs)
    h = d // 2
    s_re, s_im, r_re, r_im, o_re, o_im = (subj[:, :h].double(), subj[:, h:].double(), rel[:, :h].double(), rel[:, h:].double(),
                                          objs[:, :h].double(), objs[:, h:].double())
    ref = (s_re * r_re) @ o_re.T + (s_im * r_re) @ o_im.T + (s_re * r_im) @ o_im.T - (s_im * r_im) @ o_re.T   # :206-209
    err = (out.double() - ref).abs().max().item()
    assert err <= 4e-6 * float((subj.norm(dim=1) * rel.abs().max(dim=1).values).max() * objs.norm(dim=1).max()) * 2, err


@pytest.mark.parametrize("graph", [False, True])
def test_two_phase_epoch_keyed_regime_switches_like_the_reference(K, kats, graph):
    """openkge/trainer.py:295-299 counts the step and refreshes the epoch length BEFORE it updates the regime, and
    epoch = floor(steps / (len(loader) + 1)) + 1 (:173-175): with 3 batches per epoch the phase keyed ``epoch: 2`` starts
    at step 4, not earlier. The CUDA-graph step bakes the learning rate in and must be re-captured at the switch."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    batches = list(train.get_loader(shuffle=True, drop_last=True, seed=3))[:3]
    results = {}
    for mode in ((False, True) if graph else (False,)):
        torch.manual_seed(9)
        model = _make_model("LookupComplexRelationModel", sizes).cuda()
        args = {"optimization_config": [[{"epoch": 0, "optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10},
                                         {"epoch": 2, "lr": 0.03}]],
                "lr_scheduler_config": [None], "bce_label_smoothing": 0.0, "grad_clip": 0, "cuda_graph": mode}
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        regime = trainer.optimizers[0]
        lrs = []
        for _ in range(3):
            for b in batches:                              # one batch per call: the lr in force for exactly that step
                class One(list):
                    def __len__(self):
                        return 3                           # ... of a 3-batch epoch
                trainer.train_epoch(One([b]))
                lrs.append(regime.optimizer.param_groups[0]["lr"])
        assert trainer.training_steps == 9
        assert lrs == [0.3] * 3 + [0.03] * 6, lrs          # epoch(3) = 1, epoch(4) = 2
        assert regime.current_optimization_config_phase == 1
        assert regime.optimizer.param_groups[0]["eps"] == 1e-8          # inherited from the regime's initial Adam (:29, :143-145)
        results[mode] = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    if graph:
        for k in results[False]:
            _assert_same_trained_tensor(results[True][k], results[False][k], 0.3, 0.9, k)


@pytest.mark.parametrize("fused", [False, True])
def test_pad_row_never_receives_a_lookup_gradient(K, kats, fused):
    """nn.Embedding(padding_idx=0) (openkge/model.py:376-391): a prefix that names entity id 0 looks the PAD row up but
    never sends gradient to it -- on the plain step, the fused entity update and the graphed step alike. What still moves
    the row is the dense Adagrad's weight decay (g = wd p), exactly as in the reference."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    batch = list(train.get_loader(shuffle=True, drop_last=True, seed=3))[0]
    (po, sp) = batch[0]
    po[1][0] = 0                                           # object of the first po prefix := PAD
    sp[0][0] = 0                                           # subject of the first sp prefix := PAD
    lr, wd, eps = 0.3, 1e-3, 1e-8
    args = {"optimization_config": {"optimizer": "Adagrad", "lr": lr, "weight_decay": wd}, "lr_scheduler_config": None,
            "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": fused}
    for use_graph in (False, True):
        torch.manual_seed(9)
        model = _make_model("LookupDistmultRelationModel", sizes).cuda()
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        for o in trainer.optimizers:
            o.update(1, 1)
        p0 = model.entity_embedding.weight.data[0].double().cpu().numpy().copy()
        p5 = model.entity_embedding.weight.data[5].clone()
        if use_graph:
            step = trainer.make_graphed_step(batch, max_positives=4096, preserve_state=True)
            assert step is not None
            step(batch)
        else:
            trainer.compute_one_batch(batch, training=True, sync_loss=False)
        g = wd * p0                                        # the only gradient the PAD row sees
        want = p0 - lr * g / (np.sqrt(g * g) + eps)
        got = model.entity_embedding.weight.data[0].double().cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)
        assert not torch.equal(model.entity_embedding.weight.data[5], p5)      # ordinary rows do train


# ---------------------------------------------------------------------------------------------
# Input dropout of the candidate rows without the dropped fp32 copy (C1: input_dropout 0.4 over all 14,541 entities)
# ---------------------------------------------------------------------------------------------

def test_mask_dropout_f16_is_dropout_of_the_operand(K):
    """okge_f16_mask_dropout on the fp16 operand of x == the fp16 operand of okge_dropout(x): same Philox mask (so that the
    backward, which regenerates it, matches), kept values = the operand's values, 1 / (1 - p) in the inverse scale."""
    rng = np.random.default_rng(3)
    for rows, cols in ((97, 200), (1000, 64), (5, 4)):
        x = dev(rng.standard_normal((rows, cols)).astype(np.float32))
        op = K.quantize(x)
        p, seed, off = 0.4, 1234567, 5 << 38
        m = K.mask_dropout_f16(op, p, seed, off)
        ref = K.dropout(x, p, seed, off)
        assert torch.equal(m.dense() == 0, ref == 0)                            # the same mask
        assert 0.3 < float((ref == 0).float().mean()) < 0.5
        np.testing.assert_allclose(m.dense().cpu().numpy(), (op.dense() * (ref != 0) / (1 - p)).cpu().numpy(), rtol=1e-6)
        np.testing.assert_allclose(m.dense().cpu().numpy(), ref.cpu().numpy(), rtol=1e-3, atol=1e-3 * float(x.abs().max()))
        step = torch.tensor(3, dtype=torch.int64, device="cuda")                # replayed launches: device-side step counter
        assert torch.equal(K.mask_dropout_f16(op, p, seed, off, step).dense() == 0, K.dropout(x, p, seed, off, step) == 0)
        assert not torch.equal(K.mask_dropout_f16(op, p, seed, off, step).dense() == 0, ref == 0)


@pytest.mark.parametrize("M,N,Kd,n_ids", [(1000, 200, 130, 40), (300, 512, 64, 0), (700, 200, 2100, 20)])
def test_gemm_adagrad_dropout_matches_unfused(K, M, N, Kd, n_ids):
    """okge_gemm_adagrad_dropout == contraction -> okge_dropout of the gradient -> extra rows -> okge_adagrad_dense: the
    fused epilogue regenerates the mask of the flattened [M, N] gradient (K >= 2,048: the deep-ring instantiation)."""
    rng = np.random.default_rng(M + N + Kd)
    dS = rng.standard_normal((Kd, M)).astype(np.float32)
    q = rng.standard_normal((Kd, N)).astype(np.float32)
    a, b = _k_panels(K, dS).T, K.ColMajor(K.quantize(dev(q)))
    p0 = rng.standard_normal((M, N)).astype(np.float32)
    scale = dev(np.array([0.37], np.float32))
    ids = rng.integers(0, M, n_ids).astype(np.int32)
    rows = rng.standard_normal((n_ids, N)).astype(np.float32)
    slot_map = torch.full((M,), -1, dtype=torch.int32, device="cuda")
    p_f, G_f = dev(p0), torch.zeros(M, N, device="cuda")
    p_u, G_u = dev(p0), torch.zeros(M, N, device="cuda")
    step_dev = torch.zeros((), dtype=torch.int64, device="cuda")
    for step in range(2):
        spec = (0.4, 99, 7 << 38, step_dev)
        g = K.gemm_nt(a, b, alpha_dev=scale, splits=1).contiguous()
        K.dropout(g, *spec, out=g)
        if n_ids:
            K.scatter_add_rows(dev(rows), dev(ids), g)
        K.adagrad_dense(p_u, g, G_u, 0.3, 1e-8, 1e-10)
        extra = emap = None
        if n_ids:
            extra = torch.zeros(n_ids, N, device="cuda")
            K.row_slots_build(dev(ids), slot_map)
            K.row_slots_accumulate(dev(rows), dev(ids), slot_map, extra)
            emap = slot_map
        K.gemm_adagrad(a, b, p_f, G_f, 0.3, 1e-8, 1e-10, alpha_dev=scale, extra_map=emap, extra=extra, dropout=spec)
        if n_ids:
            K.row_slots_clear(dev(ids), slot_map)
        step_dev.add_(1)
    if n_ids == 0:        # elements dropped in both steps saw only the weight-decay term: 0.4^2 = 16 % of them
        assert 0.12 < float((G_u < 1e-12).float().mean()) < 0.20
    np.testing.assert_allclose(G_f.cpu().numpy(), G_u.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(p_f.cpu().numpy(), p_u.cpu().numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("graph", [False, True])
def test_candidate_dropout_deferred_to_the_scoring_pass_equals_dropout_of_the_rows(K, kats, graph):
    """LookupComplex with input_dropout 0.4 (the FB15k-237 config, config/fb15k237/fb15k237-complex-kge.yaml:22-30). The
    fast path keeps the candidates as the raw table (mask on the fp16 operand, mask on the gradient tile inside the fused
    dE + Adagrad step); the plain path drops the fp32 rows like the reference (F.dropout, openkge/model.py:461-470). Same
    seed and call order = same masks: the two must follow the same trajectory."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    sizes = kats["meta/sizes"]
    meta = D.EntityRelationDatasetMeta(entities_size=int(sizes[0]), relations_size=int(sizes[1]))
    tr_idx = D.PrefixIndex(kats["data/train/seen_prefixes"], kats["data/train/seen_entities"],
                           kats["data/train/all_splits_entities"], int(sizes[0]), 2, True)
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=32, device="cuda", is_training_data=True)
    batches = list(train.get_loader(shuffle=True, drop_last=True, seed=3))[:4]
    out = {}
    for mode in ("plain", "deferred", "deferred_fused"):
        torch.manual_seed(9)
        model = _make_model("LookupComplexRelationModel", sizes, input_dropout=0.4).cuda()
        model.fuse_candidate_dropout = mode != "plain"
        args = {"optimization_config": {"optimizer": "Adagrad", "lr": 0.3, "weight_decay": 1e-10}, "lr_scheduler_config": None,
                "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": mode == "deferred_fused"}
        trainer = Trainer(args, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, train)
        trainer.model_with_loss.train()
        for o in trainer.optimizers:
            o.update(1, 1)
        step = trainer.make_graphed_step(batches[0], max_positives=4096, preserve_state=True) if graph else None
        assert (step is not None) == graph
        losses = []
        for b in batches:
            if step is not None:
                losses.append(float(step(b)))
            else:
                trainer.compute_one_batch(b, training=True, sync_loss=False)
                losses.append(float(trainer.last_loss))
        if mode != "plain":
            assert model._candidates_are_raw_table and model._candidate_dropout is not None
        out[mode] = (losses, {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()})
    # same operands, gradient applied by the fused dE + Adagrad step (mask inside the epilogue) or by the dense kernels
    np.testing.assert_allclose(out["deferred_fused"][0], out["deferred"][0], rtol=1e-5)
    for k in out["plain"][1]:
        _assert_same_trained_tensor(out["deferred_fused"][1][k], out["deferred"][1][k], 0.3, 0.9, k)
    # against the dropout of the fp32 rows: the fp16 operand is rounded before instead of after the 1 / (1 - p) (scores
    # differ by a few 1e-5 relative), which four Adagrad steps turn into a few 1e-3 of a step on individual elements
    np.testing.assert_allclose(out["deferred"][0], out["plain"][0], rtol=1e-3)
    for k in out["plain"][1]:
        d = np.abs(out["deferred"][1][k].astype(np.float64) - out["plain"][1][k])
        assert (d <= 3e-3 * 0.3).mean() >= 0.95 and d.max() <= 2.1 * 0.3, (k, float((d <= 3e-3 * 0.3).mean()), float(d.max()))
