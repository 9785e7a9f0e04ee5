"""Bring-up diagnostics for the native kernels on a real B200 (prints, never asserts).

Run on the GPU box:  python scripts/gpu_bringup.py > gpurun_out/bringup.log 2>&1
"""
import sys, time, traceback
import torch

sys.path.insert(0, ".")
from open_knowledge_graph_embeddings_b200 import kernels as K

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
print("device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def section(name, fn):
    print(f"\n=== {name}")
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def gemm_case(M, N, Kd, splits=1):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + Kd)
    a = torch.randn(M, Kd, generator=g).to(dev)
    b = torch.randn(N, Kd, generator=g).to(dev)
    ref = a.double() @ b.double().t()
    out = K.gemm_nt(a, b, splits=splits)
    torch.cuda.synchronize()
    normw = (a.norm(dim=1)[:, None] * b.norm(dim=1)[None, :]).double()
    err = ((out.double() - ref).abs() / normw).max().item()
    print(f"gemm(raw operands) M={M} N={N} K={Kd} splits={splits}: normwise max err {err:.3e}  (tf32 expected ~1e-4..1e-3) "
          f"max|ref| {ref.abs().max().item():.3f} max|out| {out.abs().max().item():.3f}")
    if err > 5e-3:
        d = (out.double() - ref).abs()
        idx = d.argmax().item()
        print("   worst at", divmod(idx, N), "out", out.flatten()[idx].item(), "ref", ref.flatten()[idx].item())
        print("   out[0,:8]", out[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        bad_rows = (d.max(dim=1).values / normw.max(dim=1).values > 5e-3).nonzero().flatten()[:16].tolist()
        bad_cols = (d.max(dim=0).values / normw.max(dim=0).values > 5e-3).nonzero().flatten()[:16].tolist()
        print("   bad rows (first 16)", bad_rows, " bad cols (first 16)", bad_cols)


def t_gemm():
    for (M, N, Kd) in [(128, 256, 32), (128, 256, 64), (128, 256, 512), (256, 512, 128), (100, 300, 200),
                       (512, 14541, 200), (37, 1000, 64), (512, 4096, 512)]:
        gemm_case(M, N, Kd)
    gemm_case(512, 512, 20000, splits=8)
    gemm_case(200, 200, 14541, splits=4)


def t_gemm_speed():
    for (M, N, Kd) in [(512, 1_000_000, 512), (4096, 312_500, 512), (512, 14541, 200)]:
        a = torch.randn(M, Kd, device=dev)
        b = torch.randn(N, Kd, device=dev)
        out = torch.empty(M, K.pad4(N), device=dev)[:, :N]
        for _ in range(2):
            K.gemm_nt(a, b, out=out, splits=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 5
        for _ in range(iters):
            K.gemm_nt(a, b, out=out, splits=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"gemm store M={M} N={N} K={Kd}: {ms:.3f} ms  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s")
        del a, b, out


def make_csr(B, N, avg, g):
    rows = []
    for b in range(B):
        k = int(torch.randint(1, 2 * avg, (1,), generator=g))
        rows.append(torch.randperm(N, generator=g)[:k].sort().values)
    ptr = torch.zeros(B + 1, dtype=torch.int32)
    ptr[1:] = torch.tensor([len(r) for r in rows]).cumsum(0)
    return ptr.to(dev), torch.cat(rows).int().to(dev), rows


def t_bce():
    g = torch.Generator().manual_seed(5)
    for (B, N, D, eps) in [(64, 1000, 64, 0.0), (300, 14541, 200, 0.0), (130, 5000, 128, 0.1)]:
        q = (torch.randn(B, D, generator=g) * 0.3).to(dev)
        e = (torch.randn(N, D, generator=g) * 0.3).to(dev)
        pp, pi, rows = make_csr(B, N, 3, g)
        y = torch.zeros(B, N, device=dev)
        for b, r in enumerate(rows):
            y[b, r.to(dev)] = 1
        y_base, y_pos = 0.0, 1.0
        if eps > 0:
            y = (y + 1.0 / N) * (1 - eps)
            y_base, y_pos = (1 - eps) / N, (1 + 1.0 / N) * (1 - eps)
        s = (q.double() @ e.double().t())
        ref_loss = (torch.nn.functional.softplus(s) - s * y.double()).sum().item()
        ref_dS = torch.sigmoid(s) - y.double()
        loss, dS, dST = K.score_bce(q, e, pp, pi, y_base, y_pos)
        torch.cuda.synchronize()
        print(f"bce B={B} N={N} D={D} eps={eps}: loss {loss.item():.6f} ref {ref_loss:.6f} rel {abs(loss.item()-ref_loss)/abs(ref_loss):.2e} "
              f"dS maxabs {(dS.dense().double()-ref_dS).abs().max().item():.2e} dST maxabs {(dST.dense().double().t()-ref_dS).abs().max().item():.2e}")


def t_lse():
    g = torch.Generator().manual_seed(6)
    for (B, N, D) in [(64, 1000, 64), (300, 14541, 200)]:
        q = (torch.randn(B, D, generator=g) * 0.5).to(dev)
        e = (torch.randn(N, D, generator=g) * 0.5).to(dev)
        pp, pi, rows = make_csr(B, N, 3, g)
        s = (q.double() @ e.double().t())
        ref_lse = torch.logsumexp(s, dim=1)
        lse, ps = K.score_lse(q, e, pp, pi)
        torch.cuda.synchronize()
        ref_ps = torch.cat([s[b, r.to(dev)] for b, r in enumerate(rows)])
        print(f"lse B={B} N={N}: max abs {(lse.double()-ref_lse).abs().max().item():.2e}  pos_score max abs {(ps.double()-ref_ps).abs().max().item():.2e}")
        w = torch.tensor([float(len(r)) for r in rows], device=dev)
        dS, dST = K.score_softmax_grad(q, e, pp, pi, lse, w)
        y = torch.zeros(B, N, device=dev, dtype=torch.float64)
        for b, r in enumerate(rows):
            y[b, r.to(dev)] = 1
        ref = w.double()[:, None] * torch.softmax(s, dim=1) - y
        print(f"   softmax grad max abs {(dS.dense().double()-ref).abs().max().item():.2e}  T {(dST.dense().double().t()-ref).abs().max().item():.2e}")


def t_rank():
    g = torch.Generator().manual_seed(7)
    B, N, D = 200, 14541, 200
    q = (torch.randn(B, D, generator=g)).to(dev)
    e = (torch.randn(N, D, generator=g)).to(dev)
    e[100] = e[50]  # exact ties
    scores = K.score_store(q, e)
    s2 = K.score_store(q, e)
    print("store deterministic:", torch.equal(scores, s2))
    ans_row = torch.arange(B, dtype=torch.int32, device=dev)
    alt = torch.randint(0, N, (B,), generator=g).int().to(dev)
    alt_ptr = torch.arange(B + 1, dtype=torch.int32, device=dev)
    fp, fi, rows = make_csr(B, N, 5, g)
    true, gr, eq = K.rank_count(scores, ans_row, alt_ptr, alt, fp, fi)
    # torch check of the same counts
    masked = scores.clone()
    for b, r in enumerate(rows):
        masked[b, r.to(dev)] = -1e8
    t = scores[torch.arange(B, device=dev), alt.long()]
    gr_ref = (t[:, None] < masked).sum(1).int()
    eq_ref = (t[:, None] == masked).sum(1).int()
    print("rank_count: true eq", torch.equal(true, t), "greater eq", torch.equal(gr, gr_ref), "equal eq", torch.equal(eq, eq_ref))
    # fused count (unmasked) vs torch on the stored matrix
    g2 = torch.zeros(B, dtype=torch.int32, device=dev)
    e2 = torch.zeros(B, dtype=torch.int32, device=dev)
    K.score_rank(q, e, t, g2, e2)
    torch.cuda.synchronize()
    print("score_rank unmasked: greater eq", torch.equal(g2, (t[:, None] < scores).sum(1).int()),
          "equal eq", torch.equal(e2, (t[:, None] == scores).sum(1).int()),
          " max diff", (g2 - (t[:, None] < scores).sum(1).int()).abs().max().item())


def t_embed():
    g = torch.Generator().manual_seed(8)
    V, D, R, L, n = 1000, 64, 500, 10, 333
    W = torch.randn(V, D, generator=g).to(dev)
    id_rows = torch.randint(0, V, (R, L), generator=g).int()
    id_rows[:, 6:] = 0
    id_rows = id_rows.to(dev)
    ids = torch.randint(0, R, (n,), generator=g).int().to(dev)
    for mode in ["sum", "mean", "max"]:
        out = K.gather_pool_fwd(W, id_rows, ids, mode)
        emb = W[id_rows[ids.long()].long()]
        if mode == "sum":
            ref = emb.sum(1)
        elif mode == "mean":
            ref = emb.sum(1) / ((id_rows[ids.long()] > 0).float().sum(1, keepdim=True) + 1e-12)
        else:
            ref = emb.max(1).values
        print(f"gather_pool_fwd {mode}: max abs {(out-ref).abs().max().item():.2e}")
        go = torch.randn(n, D, generator=g).to(dev)
        gW = torch.zeros_like(W)
        K.gather_pool_bwd(go, W, id_rows, ids, mode, gW)
        Wr = W.clone().requires_grad_(True)
        embr = torch.nn.functional.embedding(id_rows[ids.long()].long(), Wr, padding_idx=0)
        if mode == "sum":
            r = embr.sum(1)
        elif mode == "mean":
            r = embr.sum(1) / ((id_rows[ids.long()] > 0).float().sum(1, keepdim=True) + 1e-12)
        else:
            r = embr.max(1).values
        r.backward(go)
        print(f"gather_pool_bwd {mode}: max abs {(gW-Wr.grad).abs().max().item():.2e}")
    out = K.gather_rows(W, ids)
    print("gather_rows eq", torch.equal(out, W[ids.long()]))
    gt = torch.zeros_like(W)
    K.scatter_add_rows(out, ids, gt)
    ref = torch.zeros_like(W).index_add_(0, ids.long(), out)
    print("scatter_add max abs", (gt - ref).abs().max().item())
    x = torch.randn(1000, 64, device=dev)
    xt = K.transpose(x)
    print("transpose eq", torch.equal(xt, x.t()))
    d = K.dropout(x, 0.4, 123, 0)
    keep = (d != 0).float().mean().item()
    print("dropout keep frac", keep, "scale ok", torch.allclose(d[d != 0], (x / 0.6)[d != 0]))
    for kind in range(3):
        a = torch.randn(50, 64, device=dev, requires_grad=True)
        b = torch.randn(50, 64, device=dev, requires_grad=True)
        qq = K.fold_query(kind, a.detach(), b.detach())
        a1, a2 = a.chunk(2, 1); b1, b2 = b.chunk(2, 1)
        if kind == 0:
            ref = torch.cat([a1 * b1 - a2 * b2, a2 * b1 + a1 * b2], 1)
        elif kind == 1:
            ref = torch.cat([a1 * b1 + a2 * b2, a2 * b1 - a1 * b2], 1)
        else:
            ref = a * b
        gq = torch.randn_like(ref)
        ref.backward(gq)
        ga, gb = K.fold_query_bwd(kind, a.detach(), b.detach(), gq)
        print(f"fold kind {kind}: fwd {(qq-ref).abs().max().item():.2e} ga {(ga-a.grad).abs().max().item():.2e} gb {(gb-b.grad).abs().max().item():.2e}")


def t_optim():
    p = torch.randn(1000, 64, device=dev); g = torch.randn_like(p); G = torch.zeros_like(p)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adagrad([pr], lr=0.3, eps=1e-8, weight_decay=1e-10)
    for step in range(3):
        pr.grad = g.clone(); opt.step()
        K.adagrad_dense(p, g, G, 0.3, 1e-8, 1e-10)
    print("adagrad dense max abs", (p - pr.detach()).abs().max().item(), "state", (G - opt.state[pr]["sum"]).abs().max().item())
    p = torch.randn(1000, 64, device=dev); m = torch.zeros_like(p); v = torch.zeros_like(p)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-2, weight_decay=1e-6)
    for step in range(1, 4):
        pr.grad = g.clone(); opt.step()
        K.adam_dense(p, g, m, v, 1e-2, 0.9, 0.999, 1e-8, 1e-6, step)
    print("adam dense max abs", (p - pr.detach()).abs().max().item())


section("embed", t_embed)
section("optim", t_optim)
section("gemm", t_gemm)
section("bce", t_bce)
section("lse", t_lse)
section("rank", t_rank)
section("gemm speed", t_gemm_speed)
print("\nDONE")


def t_precision():
    """score_store with a TF32-rounded q (as okge_fold_query produces) and a raw table operand."""
    for (B, N, D) in [(128, 2048, 32), (256, 4096, 64), (512, 14541, 200), (512, 20000, 512)]:
        g = torch.Generator().manual_seed(B + D)
        a = torch.randn(B, D, generator=g).to(dev)
        b = torch.randn(B, D, generator=g).to(dev)
        e = (torch.randn(N, D, generator=g) * 0.3).to(dev)
        q = K.fold_query(2, a, b)
        ref = (a.double() * b.double()) @ e.double().t()
        out = K.score_store(q, e)
        nw = (q.norm(dim=1)[:, None] * e.norm(dim=1)[None, :]).double()
        err = ((out.double() - ref).abs() / nw)
        print(f"score_store B={B} N={N} D={D}: normwise max {err.max().item():.3e} mean {err.mean().item():.3e} "
              f"mean signed rel bias {(((out.double()-ref)/ref)[ref.abs() > 0.1 * ref.abs().max()]).mean().item():.3e}")


section("precision", t_precision)
