"""First contact of the fp16 tensor-core path with a B200: every operand layout and epilogue is run on small shapes and
compared with an fp64 torch contraction; prints one line per case instead of stopping at the first failure (the output
of one `gpurun` call must say which descriptor / layout is wrong). Not a test: tests/test_gpu_parity.py is."""
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, ".")
from open_knowledge_graph_embeddings_b200 import kernels as K  # noqa: E402


def nw(out, ref, a, b):
    bound = a.double().norm(dim=1)[:, None] * b.double().norm(dim=1)[None, :]
    return float(((out.double() - ref).abs() / bound.clamp_min(1e-30)).max())


def case(name, fn):
    try:
        msg = fn()
        torch.cuda.synchronize()
        print(f"[ok]   {name}: {msg}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"[FAIL] {name}: {type(ex).__name__}: {ex}", flush=True)
        traceback.print_exc(limit=2)


def main():
    torch.manual_seed(0)
    dev = "cuda"
    print(torch.cuda.get_device_name(0), flush=True)
    for (M, N, Kd) in [(128, 256, 64), (128, 256, 512), (200, 300, 136), (512, 1024, 512), (64, 64, 64), (70, 40, 9000)]:
        a = torch.randn(M, Kd, device=dev)
        b = torch.randn(N, Kd, device=dev)
        ref = a.double() @ b.double().t()
        aT, bT = a.t().contiguous(), b.t().contiguous()
        forms_a = {"row": lambda: K.quantize(a), "kpan": lambda: K.Panels.from_dense(a),
                   "col": lambda: K.ColMajor(K.quantize(aT)), "mnpan": lambda: K.Panels.from_dense(aT).T}
        forms_b = {"row": lambda: K.quantize(b), "kpan": lambda: K.Panels.from_dense(b),
                   "col": lambda: K.ColMajor(K.quantize(bT)), "mnpan": lambda: K.Panels.from_dense(bT).T}
        for na, fa in forms_a.items():
            for nb, fb in forms_b.items():
                def run(fa=fa, fb=fb):
                    out = K.gemm_nt(fa(), fb(), splits=1)
                    err = nw(out, ref, a, b)
                    assert err < 1e-3, f"normwise error {err:.3e}"
                    return f"err {err:.2e}"
                case(f"f16 gemm {M}x{N}x{Kd} A={na} B={nb}", run)
        case(f"tf32 gemm {M}x{N}x{Kd} row/row", lambda: f"err {nw(K.gemm_nt(a, b, splits=1), ref, a, b):.2e}")
        case(f"tf32 gemm {M}x{N}x{Kd} col/col",
             lambda: f"err {nw(K.gemm_nt(K.ColMajor(aT), K.ColMajor(bT), splits=1), ref, a, b):.2e}")

    B, N, D = 150, 3001, 64
    q = 0.4 * torch.randn(B, D, device=dev)
    E = 0.4 * torch.randn(N, D, device=dev)
    rng = np.random.default_rng(0)
    rows = [sorted(set(rng.integers(0, N, 4).tolist())) for _ in range(B)]
    ptr = torch.tensor(np.concatenate([[0], np.cumsum([len(r) for r in rows])]), dtype=torch.int32, device=dev)
    idx = torch.tensor(np.concatenate(rows), dtype=torch.int32, device=dev)
    s = q.double() @ E.double().t()
    y = torch.zeros_like(s)
    y[torch.arange(B, device=dev).repeat_interleave(torch.tensor([len(r) for r in rows], device=dev)), idx.long()] = 1

    def bce():
        loss, dS = K.score_bce(q, E, ptr, idx)
        ref_loss = float((torch.nn.functional.softplus(s) - s * y).sum())
        ds_err = float((dS.dense().double() - (torch.sigmoid(s) - y)).abs().max())
        assert abs(loss.item() - ref_loss) < 1e-3 * abs(ref_loss) and ds_err < 1e-3, (loss.item(), ref_loss, ds_err)
        return f"loss {loss.item():.4f} ref {ref_loss:.4f} dS err {ds_err:.2e}"
    case("score_bce", bce)

    def split():
        q2, e2 = K.quantize(q, split=True), K.quantize(E, split=True)
        e1 = nw(K.score_store(q2, e2), s, q, E)
        e3 = nw(K.score_store(q2, e2, split=True), s, q, E)
        assert e3 < 4e-6 and e1 < 1e-3, (e1, e3)
        return f"single {e1:.2e} split {e3:.2e}"
    case("score_store split precision", split)

    def lse():
        l, pos = K.score_lse(q, E, ptr, idx)
        err = float((l.double() - torch.logsumexp(s, 1)).abs().max())
        assert err < 1e-2
        return f"lse err {err:.2e}"
    case("score_lse", lse)

    def rank():
        q2, e2 = K.quantize(q), K.quantize(E)
        mat = K.score_store(q2, e2)
        thr = mat[:, 5].contiguous()
        g = torch.zeros(B, dtype=torch.int32, device=dev)
        e = torch.zeros(B, dtype=torch.int32, device=dev)
        K.score_rank(q2, e2, thr, g, e)
        assert torch.equal(g, (thr[:, None] < mat).sum(1).int()) and torch.equal(e, (thr[:, None] == mat).sum(1).int())
        return "counts bit-exact"
    case("score_rank", rank)

    def adagrad():
        M, Nn, Kd = 1000, 200, 130
        dS = torch.randn(Kd, M, device=dev)
        qq = torch.randn(Kd, Nn, device=dev)
        a, b = K.Panels.from_dense(dS).T, K.ColMajor(K.quantize(qq))
        p0 = torch.randn(M, Nn, device=dev)
        p_f, G_f, p_u, G_u = p0.clone(), torch.zeros_like(p0), p0.clone(), torch.zeros_like(p0)
        shadow = K.quantize(p_f)
        K.gemm_adagrad(a, b, p_f, G_f, 0.3, 1e-8, 1e-10, alpha=0.5, shadow=shadow)
        g = K.gemm_nt(a, b, alpha=0.5, splits=1).contiguous()
        gref = 0.5 * (dS.double().t() @ qq.double())
        K.adagrad_dense(p_u, g, G_u, 0.3, 1e-8, 1e-10)
        sh_ok = torch.equal(shadow.hi[:, :Nn], (p_f / shadow.inv_scale).to(torch.float16))
        return (f"g err {float((g.double() - gref).abs().max() / gref.abs().max()):.2e} G equal {torch.equal(G_f, G_u)} "
                f"p maxdiff {float((p_f - p_u).abs().max()):.2e} shadow ok {sh_ok}")
    case("gemm_adagrad", adagrad)


if __name__ == "__main__":
    main()
