"""N > 1 arm of bench.py: entity-sharded 1-vs-all training, one rank per GPU (torchrun), NCCL over NVLink.

Weak scaling: every rank contributes ``batch`` prefix rows to the global batch (B = batch * N_gpus) and owns
1/N_gpus of the candidate rows, so each GPU scores B x N/N_gpus pairs per step — constant work per GPU — while
the triples consumed per step grow with N_gpus. ``value`` = global triples / max-over-ranks device time.
"""
import json
import os

import numpy as np
import torch
import torch.distributed as dist


def _finish():
    """End of a rank: everything is on the host already (the JSON line is flushed). Process groups that have NCCL
    collectives inside captured CUDA graphs have been seen to hang in destroy_process_group (watchdog join), so the ranks
    synchronise once more and leave without the teardown."""
    import sys
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def multi_gpu_parity(rank, world, device):
    """On-hardware correctness of the N > 1 path (CUDA kernels + NCCL), checked before anything is timed. The sharded model
    and a single-rank model (every rank runs one on its own GPU, no communication) start from the same weights:
      1. filtered ranking of a validation batch: (greater, equal) counts BIT-EQUAL (identical fp16 operands per shard,
         integer all-reduces);
      2. two training steps on the same global batches: first loss equal to 1e-6 relative (partial sums added in another
         order), second to 1e-5;
      3. post-step shard vs the single-rank rows: within the reduced-precision tolerance of a step (5e-3 lr, 99 % of the elements within 1e-4 lr). They are not
         bit-equal: the all-reduced dQ differs in the last bit, which can flip the fp16 rounding of a few elements of the
         next step's query operand (one fp16 ulp = 5e-4 relative in a term of the gradient).
    Returns "ok" or a description of the first mismatches (identical on every rank)."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    from open_knowledge_graph_embeddings_b200.sharded import EntityShardedLookupModel, shard_bounds
    spec = S.SPECS["fb15k237"]
    tr_idx, ev_idx, meta = S.build_indexes(spec, seed=3)
    N, Dm, Bg, lr = spec.n_entities, 64, 64 * world, 0.3
    g = torch.Generator(device="cpu").manual_seed(11)
    E = (torch.randn(N, Dm, generator=g) * 0.3).to(device)
    R = (torch.randn(meta.relations_size, Dm, generator=g) * 0.3).to(device)
    lo, hi = shard_bounds(N, world, rank)
    problems = []
    for scorer in ("distmult", "complex"):
        single = EntityShardedLookupModel(E.clone(), R.clone(), N, 0, 1, scorer=scorer, lr=lr, group="local")
        shard = EntityShardedLookupModel(E[lo:hi].clone(), R.clone(), N, rank, world, scorer=scorer, lr=lr)
        rng = np.random.default_rng(5)
        eb = D.input_and_labels_to_device(ev_idx.collate(rng.integers(0, len(ev_idx), Bg)), False, device)
        t1, g1, e1 = single.eval_counts(eb)
        tN, gN, eN = shard.eval_counts(eb)
        if not (torch.equal(g1, gN) and torch.equal(e1, eN)):
            problems.append(f"{scorer}: rank counts differ in {int(((g1 != gN) | (e1 != eN)).sum())} of {g1.numel()} answers")
        for step, tol in ((0, 1e-6), (1, 1e-5)):
            b = D.input_and_labels_to_device(tr_idx.collate(rng.integers(0, len(tr_idx), Bg)), True, device)
            l1, lN = float(single.train_step(b)), float(shard.train_step(b))
            if abs(l1 - lN) > tol * abs(l1):
                problems.append(f"{scorer} step {step}: loss {lN!r} vs single-rank {l1!r}")
        if hi > lo:
            d = (shard.E - single.E[lo:hi]).abs()
            if float(d.max()) > 5e-3 * lr or float((d <= 1e-4 * lr).float().mean()) < 0.99:
                problems.append(f"{scorer}: post-step shard differs from the single-rank rows by up to {float(d.max()):.3e} "
                                f"({float((d <= 1e-4 * lr).float().mean()):.4f} of the elements within 1e-4 lr)")
    flag = torch.tensor([len(problems)], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()) == 0:
        return "ok"
    return "FAILED: " + ("; ".join(problems) if problems else "mismatch on another rank")


def run_sharded(args, rank, world, device):
    import bench as B
    from open_knowledge_graph_embeddings_b200 import _capi
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    from open_knowledge_graph_embeddings_b200.sharded import EntityShardedLookupModel, shard_bounds

    workload = args.workload or B.DEFAULT_WORKLOAD
    wl = B.WORKLOADS[workload]
    parity = multi_gpu_parity(rank, world, device)
    if rank == 0 and parity != "ok":
        print(f"[multi-GPU parity] {parity}", file=__import__("sys").stderr, flush=True)
    if "Unigram" in wl["model"]:
        return run_sharded_unigram(args, rank, world, device, workload, wl, parity)
    spec = S.SPECS[wl["spec"]]
    tr_idx, ev_idx, meta = S.build_indexes(spec, seed=1)
    N, Dm = spec.n_entities, wl["dim"]
    lo, hi = shard_bounds(N, world, rank)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    E = (torch.randn(hi - lo, Dm, generator=g) * 0.1).to(device)
    g2 = torch.Generator(device="cpu").manual_seed(7)                      # relation table replicated: same seed
    R = (torch.randn(meta.relations_size, Dm, generator=g2) * 0.1).to(device)
    model = EntityShardedLookupModel(E, R, N, rank, world, scorer="complex" if "Complex" in wl["model"] else "distmult",
                                     lr=wl["lr"], eps=1e-8, weight_decay=wl["weight_decay"])
    K, W = args.steps, args.warmup
    Bg = wl["batch"] * world
    pool = [tr_idx.collate(r, pin=True) for r in np.random.default_rng(7).integers(0, len(tr_idx), (min(K + W, 16), Bg))]     # identical on every rank (same seed)
    dev_pool = [D.input_and_labels_to_device(b, True, device, non_blocking=False) for b in pool]

    timer = B.KernelTimer()
    if rank == 0:
        _capi.set_call_hook(timer.hook)
    sampler = B.ClockSampler(device.index)

    step_fn = {"fn": model.train_step}

    def timed(batches, to_device, read_loss):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        triples, h2d = 0.0, 0
        e0.record()
        for i in range(K):
            b = batches[(W + i) % len(batches)]
            if to_device:
                h2d += D.batch_h2d_bytes(b)
                b = D.input_and_labels_to_device(b, True, device)
            loss = step_fn["fn"](b)
            if read_loss:
                loss.item()
            triples += b[2] / 2.0
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return triples, float(ms.item()), h2d

    # leg 0: per-kernel breakdown on rank 0 (CUDA events around every native call; eager launches, not timed as a whole)
    for i in range(W):
        model.train_step(dev_pool[i % len(dev_pool)])
    torch.cuda.synchronize()
    K0 = min(K, 10)
    timer.enabled = rank == 0
    for i in range(K0):
        model.train_step(dev_pool[(W + i) % len(dev_pool)])
    torch.cuda.synchronize()
    timer.enabled = False
    launches_per_step = timer.launches / K0 if rank == 0 else 0

    # The timed legs replay the step (kernels + NCCL all-reduces) as one CUDA graph per rank; every rank must take the same
    # path, so the outcome of the capture is agreed on with an all-reduce.
    graph_note = "disabled (--no-cuda-graph)"
    if not args.no_cuda_graph:
        from open_knowledge_graph_embeddings_b200.sharded import GraphedShardedStep
        ok = torch.ones(1, device=device)
        gstep = None
        try:
            gstep = GraphedShardedStep(model, Bg, max(4096, 2 * max(int(b[3].idx.numel()) for b in pool)), dev_pool[0])
        except Exception as ex:  # noqa: BLE001
            ok.zero_()
            graph_note = f"capture failed, eager launches: {type(ex).__name__}: {str(ex)[:120]}"
        torch.cuda.synchronize()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0:
            step_fn["fn"] = gstep
            graph_note = "sharded step (kernels + NCCL all-reduces) replayed as one CUDA graph per rank"
        elif gstep is not None:
            graph_note = "capture failed on another rank, eager launches"

    for i in range(W):
        step_fn["fn"](dev_pool[i % len(dev_pool)])
    if rank == 0:
        sampler.start()
    triples, ms_total, _ = timed(dev_pool, to_device=False, read_loss=False)
    for i in range(2):
        step_fn["fn"](D.input_and_labels_to_device(pool[i], True, device)).item()
    triples2, ms_e2e, h2d = timed(pool, to_device=True, read_loss=True)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = B.load_peaks()
        agg = timer.summary()
        roof = B.roofline_of(agg, peaks, {}, workload)
        if roof:
            for k, v in roof["breakdown"].items():
                v["ms_per_step"] = round(v["total_ms"] / K0, 4)
            roof["breakdown_note"] = f"CUDA events around every native call over {K0} eagerly launched steps on rank 0"
        out = {"metric": B.METRIC, "value": round(triples / (ms_total / 1e3), 1), "unit": B.UNIT, "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": round(ms_total / K, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f16xf16+f32acc", "data": "synthetic",
               "config": B.config_of(workload, wl, world, wl["batch"]),
               "e2e": {"value": round(triples2 / (ms_e2e / 1e3), 1), "unit": B.UNIT, "h2d_bytes_per_step": int(h2d / K),
                       "d2h_bytes_per_step": 8, "ms_per_step": round(ms_e2e / K, 4)},
               "gpu_launches": int(round(launches_per_step * K)), "cuda_graph": step_fn["fn"] is not model.train_step,
               "cuda_graph_note": graph_note, "clocks": clocks, "roofline": roof, "multi_gpu_parity": parity,
               "collectives_per_step": ["all_reduce X[B,D] f32", "all_reduce dQ[B,D] f32", "all_reduce loss f64"],
               "prefix_rows_per_sec": round(K * Bg / (ms_total / 1e3), 1)}
        print(json.dumps(out), flush=True)
    _finish()


def run_sharded_unigram(args, rank, world, device, workload, wl, parity=None):
    """Token-model workloads over N GPUs (sharded.CandidateShardedUnigramModel): C4 = batch-shared BCE training, global
    batch 4096 x N rows, candidate list partitioned; C5 = filtered evaluation of 1024 x N queries per step against the
    2.5 M mentions, pooled-embedding cache partitioned by rows. Weak scaling: the scoring work per GPU is constant."""
    import bench as B
    from open_knowledge_graph_embeddings_b200 import _capi
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    from open_knowledge_graph_embeddings_b200.sharded import CandidateShardedUnigramModel

    spec = S.SPECS[wl["spec"]]
    tr_idx, ev_idx, meta = S.build_indexes(spec, seed=1, scale=wl.get("scale", 1.0))
    Dm = wl["dim"]
    g = torch.Generator(device="cpu").manual_seed(7)                       # replicated state: same seed on every rank
    params = {"entity_embedding.weight": torch.randn(meta.entity_tokens_size, Dm, generator=g) * 0.1,
              "relation_embedding.weight": torch.randn(meta.relation_tokens_size, Dm, generator=g) * 0.1,
              "entity_token_ids": torch.from_numpy(meta.entity_token_rows), "relation_token_ids": torch.from_numpy(meta.relation_token_rows)}
    if wl["model_config"].get("normalize") == "batchnorm":
        for which in ("entity", "relation"):
            params.update({f"{which}_batchnorm.weight": torch.rand(Dm, generator=g), f"{which}_batchnorm.bias": torch.zeros(Dm),
                           f"{which}_batchnorm.running_mean": torch.zeros(Dm), f"{which}_batchnorm.running_var": torch.ones(Dm),
                           f"{which}_batchnorm.num_batches_tracked": torch.zeros((), dtype=torch.int64)})
    params = {k: v.to(device) for k, v in params.items()}
    model = CandidateShardedUnigramModel(params, spec.n_entities, rank, world, scorer="complex" if "Complex" in wl["model"] else "distmult",
                                         lr=wl["lr"], eps=1e-8, weight_decay=wl["weight_decay"])
    K, W = args.steps, args.warmup
    Bg = wl["batch"] * world
    eval_only = wl.get("eval_only", False)
    np.random.seed(1)
    rng = np.random.default_rng(7)
    n_pool = min(K + W, 12)
    if eval_only:
        pool = [ev_idx.collate(rng.integers(0, len(ev_idx), Bg), pin=True) for _ in range(n_pool)]
    elif wl.get("shared"):
        pool = [D.collate_shared(tr_idx, rng.integers(0, len(tr_idx), Bg), wl.get("min_size_batch_labels", -1) * world, pin=True)
                for _ in range(n_pool)]
    else:
        pool = [tr_idx.collate(rng.integers(0, len(tr_idx), Bg), pin=True) for _ in range(n_pool)]
    dev_pool = [D.input_and_labels_to_device(b, not eval_only, device, non_blocking=False) for b in pool]
    timer = B.KernelTimer()
    if rank == 0:
        _capi.set_call_hook(timer.hook)
    sampler = B.ClockSampler(device.index)

    def one(b):
        if eval_only:
            _, greater, equal = model.eval_counts(b)
            return float(greater.numel()), (greater, equal)
        return b[2] / 2.0, model.train_step(b)

    def timed(batches, host):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        units, h2d = 0.0, 0
        e0.record()
        for i in range(K):
            b = batches[(W + i) % len(batches)]
            if host:
                h2d += D.batch_h2d_bytes(b)
                b = D.input_and_labels_to_device(b, not eval_only, device)
            u, res = one(b)
            units += u
            if host:                                                       # the step's result reaches the host
                (D.metrics_from_counts(*res) if eval_only else res.item())
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return units, float(ms.item()), h2d

    with torch.no_grad():
        for i in range(W):
            one(dev_pool[i % len(dev_pool)])
        if rank == 0:
            sampler.start()
        timer.enabled = rank == 0
        units, ms_total, _ = timed(dev_pool, host=False)
        timer.enabled = False
        units2, ms_e2e, h2d = timed(pool, host=True)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        roof = B.roofline_of(timer.summary(), B.load_peaks(), {}, workload)
        if roof:
            for k, v in roof["breakdown"].items():
                v["ms_per_step"] = round(v["total_ms"] / K, 4)
        metric, unit = (("filtered_eval_queries_per_sec", "queries/s") if eval_only else (B.METRIC, B.UNIT))
        cfg = B.config_of(workload, wl, world, wl["batch"])
        cfg["parallelism"] = f"candidate-sharded x{world} (token tables replicated)"
        out = {"metric": metric, "value": round(units / (ms_total / 1e3), 1), "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f16xf16+f32acc", "data": "synthetic", "config": cfg,
               "e2e": {"value": round(units2 / (ms_e2e / 1e3), 1), "unit": unit, "h2d_bytes_per_step": int(h2d / K),
                       "d2h_bytes_per_step": 52 if eval_only else 8, "ms_per_step": round(ms_e2e / K, 4)},
               "gpu_launches": timer.launches, "clocks": clocks, "roofline": roof, "multi_gpu_parity": parity,
               "collectives_per_step": (["all_reduce(max) true scores f32 [Q]", "all_reduce(sum) greater/equal int32 [Q]"] if eval_only else
                                        ["all_reduce BN sums f64 [2D+1] (fwd) + [2D] (bwd)", "all_reduce dQ[B,D] f32", "all_reduce loss f64",
                                         "all_reduce token-table grad f32 [V,D]"]),
               "prefix_rows_per_sec": round(K * Bg / (ms_total / 1e3), 1)}
        print(json.dumps(out), flush=True)
    _finish()
