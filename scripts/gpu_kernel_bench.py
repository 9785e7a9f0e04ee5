"""Times the tensor-core kernels of one training step in isolation at a BASELINE shape (default C3: N = 10^6, D = 512,
B = 512) with CUDA events: forward + loss (okge_score_bce), dQ (okge_gemm_f16_nt, split-K), fused dE + Adagrad
(okge_gemm_adagrad), single-pass evaluation (okge_score_bce_rank, single and split precision). Operands are larger than
L2, so every iteration streams from HBM. Used for A/B runs of kernel changes and as the command profiled with ncu."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from open_knowledge_graph_embeddings_b200 import kernels as K  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--D", type=int, default=512)
    ap.add_argument("--B", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    N, D, B = args.N, args.D, args.B
    torch.manual_seed(0)
    dev = "cuda"
    E = torch.randn(N, D, device=dev) * 0.1
    G = torch.zeros_like(E)
    q = torch.randn(B, D, device=dev) * 0.1
    ptr = torch.arange(0, 2 * B + 1, 2, dtype=torch.int32, device=dev)
    idx = torch.sort(torch.randint(0, N, (B, 2), device=dev), dim=1).values.reshape(-1).to(torch.int32)
    e16, q16 = K.quantize(E, split=True), K.quantize(q, split=True)
    e1, q1 = e16.without_lo(), q16.without_lo()
    scale = torch.tensor([1.0 / (B * N)], device=dev)
    out = {"shape": dict(N=N, D=D, B=B)}
    want = set(args.only.split(",")) if args.only else None

    def on(name):
        return want is None or name in want

    loss, dS = K.score_bce(q1, e1, ptr, idx)
    if on("bce"):
        ms = timed(lambda: K.score_bce(q1, e1, ptr, idx), args.iters)
        out["score_bce"] = dict(ms=ms, tflops=2.0 * B * N * D / ms / 1e9, gbs=(2.0 * N * D + 2.0 * B * N) / ms / 1e6)
        ms = timed(lambda: K.score_bce(q1, e1, ptr, idx, want_dS=False), args.iters)
        out["score_bce_loss_only"] = dict(ms=ms, tflops=2.0 * B * N * D / ms / 1e9)
    if on("dq"):
        ms = timed(lambda: K.gemm_nt(dS, K.ColMajor(e1), alpha_dev=scale), args.iters)
        out["dQ"] = dict(ms=ms, tflops=2.0 * B * N * D / ms / 1e9, gbs=(2.0 * N * D + 2.0 * B * N) / ms / 1e6)
    if on("adagrad"):
        ms = timed(lambda: K.gemm_adagrad(dS.T, K.ColMajor(q1), E, G, 0.3, 1e-8, 1e-10, alpha_dev=scale, shadow=e1), args.iters)
        out["gemm_adagrad"] = dict(ms=ms, gbs=(18.0 * N * D + 2.0 * B * N) / ms / 1e6, tflops=2.0 * B * N * D / ms / 1e9)
        ms = timed(lambda: K.gemm_adagrad(dS.T, K.ColMajor(q1), E, G, 0.3, 1e-8, 1e-10, alpha_dev=scale), args.iters)
        out["gemm_adagrad_no_shadow"] = dict(ms=ms, gbs=(16.0 * N * D + 2.0 * B * N) / ms / 1e6)
        g = torch.zeros_like(E)
        ms = timed(lambda: K.adagrad_dense(E, g, G, 0.3, 1e-8, 1e-10), args.iters)
        out["adagrad_dense"] = dict(ms=ms, gbs=20.0 * N * D / ms / 1e6)
        del g
    if on("de"):   # the unfused large-batch update route: plain dE contraction + dense Adagrad + re-quantised shadow
        dE = torch.zeros_like(E)
        ms = timed(lambda: K.gemm_nt(dS.T, K.ColMajor(q1), alpha_dev=scale, out=dE, splits=1), args.iters)
        out["dE"] = dict(ms=ms, tflops=2.0 * B * N * D / ms / 1e9, gbs=(4.0 * N * D + 2.0 * B * N) / ms / 1e6)
        ms = timed(lambda: K.adagrad_dense(E, dE, G, 0.3, 1e-8, 1e-10), args.iters)
        out["adagrad_dense_after_dE"] = dict(ms=ms, gbs=20.0 * N * D / ms / 1e6)
        ms = timed(lambda: K.quantize(E, out=e1), args.iters)
        out["requantize_shadow"] = dict(ms=ms)
        del dE
    if on("eval"):
        Bq = B + 64
        qe = K.quantize(torch.randn(Bq, D, device=dev) * 0.1, split=True)
        thr = torch.zeros(Bq * 4, device=dev)
        g4 = torch.zeros(Bq * 4, dtype=torch.int32, device=dev)
        e4 = torch.zeros(Bq * 4, dtype=torch.int32, device=dev)
        lo = torch.zeros(1, dtype=torch.float64, device=dev)
        for split in (False, True):
            for slots in (1, 4):
                ms = timed(lambda: K.score_bce_rank(qe, e16, ptr, idx, 0.0, 1.0, thr, g4, e4, lo, extra_rows=64, split=split,
                                                    slots=slots), args.iters)
                terms = 3 if split else 1
                out[f"score_bce_rank_{'split' if split else 'single'}_slots{slots}"] = dict(
                    ms=ms, tflops=2.0 * terms * Bq * N * D / ms / 1e9)
        thr1 = torch.zeros(Bq, device=dev)
        g1 = torch.zeros(Bq, dtype=torch.int32, device=dev)
        for split in (False, True):
            ms = timed(lambda: K.score_rank(qe, e16, thr1, g1, g1.clone(), split=split), args.iters)
            terms = 3 if split else 1
            out[f"score_rank_{'split' if split else 'single'}"] = dict(ms=ms, tflops=2.0 * terms * Bq * N * D / ms / 1e9)
    if on("quantize"):
        ms = timed(lambda: K.quantize(E, out=e1), args.iters)
        out["quantize_table"] = dict(ms=ms, gbs=6.0 * N * D / ms / 1e6 + 4.0 * N * D / ms / 1e6)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
