"""Isolated timings of the hot kernels at the C3 shape (run on a B200 through gpurun, optionally under ncu).

    python scripts/gpu_kernel_bench.py [kernel ...]     kernels: bce dq de adagrad fused rank bcerank store pool

Each kernel runs on operands larger than L2 (N = 10^6 candidates, D = 512, B = 512); CUDA events, 3 warm-ups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_knowledge_graph_embeddings_b200 import kernels as K  # noqa: E402


def timed(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def pool_bench(iters):
    """Token gather + pooling at the OLPBench shape (C4 / C5): 2.5 M mention rows x 10 token slots x D = 512 from a
    200 k-token table (410 MB, does not fit L2), Zipf-ish token popularity; forward (the eval cache build) and backward."""
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    import numpy as np
    dev = torch.device("cuda")
    n, L, D, V = int(os.environ.get("OKGE_POOL_ROWS", 2_500_000)), 10, 512, 200_000
    rows = S.token_rows(np.random.default_rng(0), n, V, 4.8)
    id_rows = torch.from_numpy(rows).to(torch.int32).to(dev)
    W = torch.randn(V + 4, D, device=dev) * 0.1
    for mode in ("sum", "mean", "max"):
        out = K.gather_pool_fwd(W, id_rows, None, mode, 2, n)
        ms = timed(lambda: K.gather_pool_fwd(W, id_rows, None, mode, 2, n), iters)
        alg = n * (4.0 * L + 4.0 * L * D + 4.0 * D)
        print(f"gather_pool_fwd[{mode:4s}] {ms:8.3f} ms  {alg / ms / 1e6:8.1f} GB/s algorithmic ({alg / 1e9:.1f} GB: ids + every gathered token row + output)"
              f"  |  output-only {n * 4.0 * D / ms / 1e6:7.1f} GB/s")
    g = torch.randn(n, D, device=dev)
    gw = torch.zeros_like(W)
    ms = timed(lambda: K.gather_pool_bwd(g, W, id_rows, None, "sum", gw, 2), iters)
    alg = n * (4.0 * D + 4.0 * L + 8.0 * L * D)
    print(f"gather_pool_bwd[sum ] {ms:8.3f} ms  {alg / ms / 1e6:8.1f} GB/s algorithmic ({alg / 1e9:.1f} GB: grad rows + ids + RMW of every token slot)"
          f"  |  input-only {n * 4.0 * D / ms / 1e6:7.1f} GB/s")


def main():
    which = sys.argv[1:] or ["bce", "dq", "de", "adagrad", "fused", "rank"]
    if "pool" in which:
        pool_bench(int(os.environ.get("OKGE_ITERS", 5)))
        which = [w for w in which if w != "pool"]
        if not which:
            return
    N = int(os.environ.get("OKGE_N", 1_000_000))
    D = int(os.environ.get("OKGE_D", 512))
    B = int(os.environ.get("OKGE_B", 512))
    iters = int(os.environ.get("OKGE_ITERS", 5))
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    E = torch.randn(N, D, device=dev, generator=g) * 0.1
    q = K.fold_query(K.FOLD_DISTMULT, torch.randn(B, D, device=dev, generator=g) * 0.1,
                     torch.randn(B, D, device=dev, generator=g) * 0.1 + 1.0)
    ptr = torch.arange(0, B + 1, dtype=torch.int32, device=dev)
    idx = torch.randint(0, N, (B,), dtype=torch.int32, device=dev)
    loss, dS, _ = K.score_bce(q, E, ptr, idx, want_dST=False)
    scale = torch.tensor([1.0 / (B * N)], device=dev)
    G = torch.zeros_like(E)
    flops = 2.0 * B * N * D
    for name in which:
        if name == "bce":
            ms = timed(lambda: K.score_bce(q, E, ptr, idx, want_dST=False), iters)
            print(f"score_bce          {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
        elif name == "dq":
            ms = timed(lambda: K.gemm_nt(dS, K.ColMajor(E), alpha_dev=scale), iters)
            print(f"dQ = dS E          {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
        elif name == "de":
            out = torch.empty_like(E)
            ms = timed(lambda: K.gemm_nt(dS.T, K.ColMajor(q), alpha_dev=scale, out=out, splits=1), iters)
            print(f"dE = dS^T Q        {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s  {(4.0 * B * N + 4.0 * N * D) / ms / 1e6:8.1f} GB/s")
        elif name == "adagrad":
            grad = torch.randn_like(E)
            ms = timed(lambda: K.adagrad_dense(E, grad, G, 0.3, 1e-8, 1e-10), iters)
            print(f"adagrad_dense      {ms:8.3f} ms  {20.0 * N * D / ms / 1e6:8.1f} GB/s")
        elif name == "fused":
            ms = timed(lambda: K.gemm_adagrad(dS.T, K.ColMajor(q), E, G, 0.3, 1e-8, 1e-10, alpha_dev=scale), iters)
            print(f"gemm_adagrad       {ms:8.3f} ms  {(16.0 * N * D + 4.0 * B * N) / ms / 1e6:8.1f} GB/s")
        elif name == "rank":
            thr = torch.zeros(B, device=dev)
            gr = torch.zeros(B, dtype=torch.int32, device=dev)
            eq = torch.zeros(B, dtype=torch.int32, device=dev)
            ms = timed(lambda: K.score_rank(q, E, thr, gr, eq), iters)
            print(f"score_rank         {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
        elif name == "bcerank":
            # evaluation step: loss-only BCE pass + count pass (one row per ranked answer, ~1.1 per prefix) vs the single pass
            thr4 = torch.full((B, 4), float("inf"), device=dev)
            thr4[:, 0] = 0.0
            thr4[: B // 8, 1] = 0.1
            g4 = torch.zeros((B, 4), dtype=torch.int32, device=dev)
            e4 = torch.zeros((B, 4), dtype=torch.int32, device=dev)
            out = torch.zeros(1, dtype=torch.float64, device=dev)
            ms = timed(lambda: K.score_bce_rank(q, E, ptr, idx, 0.0, 1.0, thr4, g4, e4, out), iters)
            print(f"score_bce_rank     {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
            nq = B + B // 8
            qx = q[torch.arange(nq, device=dev) % B].contiguous()
            thr = torch.zeros(nq, device=dev)
            gr = torch.zeros(nq, dtype=torch.int32, device=dev)
            eq = torch.zeros(nq, dtype=torch.int32, device=dev)
            ms1 = timed(lambda: K.score_bce(q, E, ptr, idx, want_dS=False, want_dST=False), iters)
            ms2 = timed(lambda: K.score_rank(qx, E, thr, gr, eq), iters)
            print(f"two passes         {ms1 + ms2:8.3f} ms  (loss-only score_bce {ms1:.3f} + score_rank over {nq} rows {ms2:.3f})")
        elif name == "store":
            Ns = min(N, 100_000)
            ms = timed(lambda: K.score_store(q, E[:Ns]), iters)
            print(f"score_store[{Ns}] {ms:8.3f} ms  {2.0 * B * Ns * D / ms / 1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
