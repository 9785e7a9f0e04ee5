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


def _trainer_for(model, train, valid, lr, wd, sharded, fused=True):
    from open_knowledge_graph_embeddings_b200.trainer import Trainer
    targs = {"optimization_config": {"optimizer": "Adagrad", "lr": lr, "weight_decay": wd}, "lr_scheduler_config": None,
             "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": fused, "entity_sharding": sharded}
    t = Trainer(targs, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)
    for o in t.optimizers:
        o.update(1, 1)
    return t


def _eval_counts(trainer, valid, batch):
    """(loss, greater, equal) of one evaluation batch through AddLossModule.forward + rank_answers."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    inputs, nl, nm, labels, label_ids, filt, shared = batch
    trainer.model.eval()
    with torch.no_grad():
        loss, _, pred = trainer.model_with_loss(inputs=inputs, labels=labels, batch_shared_entities=shared,
                                                use_batch_shared_entities=False, epoch=1,
                                                input_style_triple_or_prefix=valid.input_style)
        _, g, e, _ = D.rank_answers(filt, label_ids, pred)
    trainer.model.train()
    return float(loss), g, e


SECONDARY = ("c4_olpbench_unigram", "c5_olpbench_eval")


def main_sharded(args, rank, world, device):
    """The N > 1 arm: the requested (default: BASELINE's headline) workload, and -- for the default run only -- short runs
    of the OLPBench-shaped workloads in the same job, summarised under "secondary" of the ONE JSON line rank 0 prints."""
    import bench as B
    out = run_sharded(args, rank, world, device)
    if args.workload is None and not args.no_secondary:
        parity = out.get("multi_gpu_parity") if out is not None else "see rank 0"
        short = type(args)(**vars(args))
        short.steps, short.warmup = min(args.steps, 10), 3
        secondary = {}
        # a secondary workload that hangs (a collective one rank never reaches) must not cost the headline line
        import threading

        def give_up():
            if rank == 0:
                out["secondary"] = dict(secondary, error="timed out after 300 s")
                print(json.dumps(out), flush=True)
            os._exit(0)
        watchdog = threading.Timer(300.0, give_up)
        watchdog.daemon = True
        watchdog.start()
        for w in SECONDARY:
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            try:
                sub = run_sharded(short, rank, world, device, workload=w, parity=parity)
                if rank == 0:
                    secondary[w] = B.compact_line(sub)
            except Exception as ex:  # noqa: BLE001  (a secondary workload must not cost the headline line)
                secondary[w] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        watchdog.cancel()
        if rank == 0:
            out["secondary"] = secondary
    if rank == 0:
        print(json.dumps(out), flush=True)
    _finish()


def multi_gpu_parity(rank, world, device):
    """On-hardware correctness of the N > 1 path (CUDA kernels + NCCL), checked before anything is timed, THROUGH THE PUBLIC
    API: the same ``Models`` class under the same ``Trainer``, once with its entity table partitioned over the ranks
    (``model.shard_entities``, what ``Trainer`` does in a multi-rank job) and once unsharded (every rank runs that one on
    its own GPU, no communication), from the same weights:
      1. filtered ranking of a validation batch: (greater, equal) counts BIT-EQUAL (identical fp16 operands per block,
         integer all-reduces), loss equal to 1e-6 relative;
      2. two training steps on the same global batches: first loss equal to 1e-6 relative (partial sums added in another
         order), second to 1e-5;
      3. post-step block vs the unsharded model's rows: 99 % of the elements within 1e-4 lr, none further than 2 % of a
         learning-rate step. They are not bit-equal: the all-reduced dQ differs in the last bit, which can flip the fp16
         rounding of a few elements of the next step's query operand (one fp16 ulp = 5e-4 relative in a term of the
         gradient), and Adagrad's first steps (update = lr g / sqrt(sum g^2)) are ill-conditioned where g is tiny.
    Returns "ok" or a description of the first mismatches (identical on every rank)."""
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200 import synthetic as S
    from open_knowledge_graph_embeddings_b200.model import Models
    spec = S.SPECS["fb15k237"]
    tr_idx, ev_idx, meta = S.build_indexes(spec, seed=3)
    Dm, Bg, lr = 64, 64 * world, 0.3
    train = D.OneToNMentionRelationDataset(tr_idx, meta, batch_size=Bg, device=device, is_training_data=True)
    valid = D.OneToNMentionRelationDataset(ev_idx, meta, batch_size=Bg, device=device, is_training_data=False)
    problems = []
    for name in ("LookupDistmultRelationModel", "LookupComplexRelationModel"):
        models = []
        for _ in range(2):
            torch.manual_seed(11)                               # identical initial weights, on every rank
            models.append(getattr(Models, name)(entity_slot_size=Dm, init_std=0.3, train_data=meta).to(device))
        single = _trainer_for(models[0], train, valid, lr, 1e-10, sharded=False)
        shard = _trainer_for(models[1], train, valid, lr, 1e-10, sharded=True)
        sh = models[1]._shard
        if world > 1 and sh is None:
            problems.append(f"{name}: Trainer did not shard the entity table")
            continue
        rng = np.random.default_rng(5)
        eb = D.input_and_labels_to_device(ev_idx.collate(rng.integers(0, len(ev_idx), Bg)), False, device)
        l1, g1, e1 = _eval_counts(single, valid, eb)
        lN, gN, eN = _eval_counts(shard, valid, eb)
        if not (torch.equal(g1, gN) and torch.equal(e1, eN)):
            problems.append(f"{name}: rank counts differ in {int(((g1 != gN) | (e1 != eN)).sum())} of {g1.numel()} answers")
        if abs(l1 - lN) > 1e-6 * abs(l1):
            problems.append(f"{name}: eval loss {lN!r} vs single-rank {l1!r}")
        for step, tol in ((0, 1e-6), (1, 1e-5)):
            b = D.input_and_labels_to_device(tr_idx.collate(rng.integers(0, len(tr_idx), Bg)), True, device)
            single.compute_one_batch(b, training=True, sync_loss=False)
            shard.compute_one_batch(b, training=True, sync_loss=False)
            l1, lN = float(single.last_loss), float(shard.last_loss)
            if abs(l1 - lN) > tol * abs(l1):
                problems.append(f"{name} step {step}: loss {lN!r} vs single-rank {l1!r}")
        full, block = models[0].entity_embedding.weight.data, models[1].entity_embedding.weight.data
        lo, hi = (sh.lo, sh.hi) if sh is not None else (0, full.size(0) - 2)
        if hi > lo:
            d = (block[2:] - full[2 + lo:2 + hi]).abs()
            if float(d.max()) > 2e-2 * lr or float((d <= 1e-4 * lr).float().mean()) < 0.99:
                problems.append(f"{name}: post-step block differs from the single-rank rows by up to {float(d.max()):.3e} "
                                f"({float((d <= 1e-4 * lr).float().mean()):.4f} of the elements within 1e-4 lr)")
        dr = (models[1].relation_embedding.weight.data - models[0].relation_embedding.weight.data).abs().max()
        if float(dr) > 5e-3 * lr:
            problems.append(f"{name}: replicated relation table differs from the single-rank one by {float(dr):.3e}")
    everyone = [None] * world
    dist.all_gather_object(everyone, problems)
    found = [f"rank {r}: {p}" for r, ps in enumerate(everyone) for p in ps]
    return "ok" if not found else "FAILED: " + "; ".join(found[:6])


def run_sharded(args, rank, world, device, workload=None, parity=None):
    """Lookup workloads over N GPUs through the public API: ``Models.<name>`` under ``Trainer`` in a torch.distributed job
    (the Trainer partitions the entity table, ``model.shard_entities``), the step replayed as one CUDA graph per rank
    (``Trainer.make_graphed_step``: kernels + NCCL all-reduces)."""
    import bench as B
    from open_knowledge_graph_embeddings_b200 import _capi
    from open_knowledge_graph_embeddings_b200 import dataset as D

    workload = workload or args.workload or B.DEFAULT_WORKLOAD
    wl = B.WORKLOADS[workload]
    if parity is None:
        parity = multi_gpu_parity(rank, world, device)
        if rank == 0 and parity != "ok":
            print(f"[multi-GPU parity] {parity}", file=__import__("sys").stderr, flush=True)
    if wl.get("shared") and not args.sharded_engine:
        return run_data_parallel(args, rank, world, device, workload, wl, parity)
    if "Unigram" in wl["model"]:
        return run_sharded_unigram(args, rank, world, device, workload, wl, parity)
    Bg = wl["batch"] * world
    wl, spec, model, train, valid = B.build_workload(workload, device, world, rank, batch=Bg)   # same seed: same init everywhere
    trainer = _trainer_for(model, train, valid, wl["lr"], wl["weight_decay"], sharded=True, fused=not args.unfused_update)
    torch.cuda.empty_cache()                                   # the unsharded initial table
    trainer.model_with_loss.train()
    K, W = args.steps, args.warmup
    tr_idx = train.index
    pool = [tr_idx.collate(r, pin=True) for r in np.random.default_rng(7).integers(0, len(tr_idx), (min(K + W, 16), Bg))]     # identical on every rank (same seed)
    dev_pool = [D.input_and_labels_to_device(b, True, device, non_blocking=False) for b in pool]

    timer = B.KernelTimer()
    if rank == 0:
        _capi.set_call_hook(timer.hook)
    sampler = B.ClockSampler(device.index)

    def eager(b, sync_loss=False):
        trainer.compute_one_batch(b, training=True, sync_loss=False)
        if sync_loss:
            trainer.last_loss.item()

    step_fn = {"fn": eager}

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
            step_fn["fn"](b, read_loss)
            triples += b[2] / 2.0
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return triples, float(ms.item()), h2d

    # leg 0: per-kernel breakdown on rank 0 (CUDA events around every native call; eager launches, not timed as a whole)
    for i in range(W):
        eager(dev_pool[i % len(dev_pool)])
    torch.cuda.synchronize()
    K0 = min(K, 10)
    timer.enabled = rank == 0
    for i in range(K0):
        eager(dev_pool[(W + i) % len(dev_pool)])
    torch.cuda.synchronize()
    timer.enabled = False
    launches_per_step = timer.launches / K0 if rank == 0 else 0

    # The timed legs replay the step (kernels + NCCL all-reduces) as one CUDA graph per rank; every rank must take the same
    # path, so the outcome of the capture is agreed on with an all-reduce.
    graph_note = "disabled (--no-cuda-graph)"
    graphed = False
    if not args.no_cuda_graph:
        ok = torch.ones(1, device=device)
        gstep = None
        try:
            gstep = trainer.make_graphed_step(dev_pool[0], max(4096, 2 * max(int(b[3].idx.numel()) for b in pool)))
            if gstep is None:
                ok.zero_()
                graph_note = "configuration not capturable, eager launches"
        except Exception as ex:  # noqa: BLE001
            ok.zero_()
            graph_note = f"capture failed, eager launches: {type(ex).__name__}: {str(ex)[:120]}"
        torch.cuda.synchronize()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0:
            step_fn["fn"] = lambda b, s: gstep.step(b, s)
            graphed = True
            graph_note = "sharded step (kernels + NCCL all-reduces) replayed as one CUDA graph per rank"
        elif gstep is not None:
            graph_note = "capture failed on another rank, eager launches"

    for i in range(W):
        step_fn["fn"](dev_pool[i % len(dev_pool)], False)
    if rank == 0:
        sampler.start()
    triples, ms_total, _ = timed(dev_pool, to_device=False, read_loss=False)
    for i in range(2):
        step_fn["fn"](D.input_and_labels_to_device(pool[i], True, device), True)
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
        cfg = B.config_of(workload, wl, world, wl["batch"])
        cfg["api"] = "Models.%s under Trainer; Trainer shards the entity table over the ranks (model.shard_entities)" % wl["model"]
        out = {"metric": B.METRIC, "value": round(triples / (ms_total / 1e3), 1), "unit": B.UNIT, "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": round(ms_total / K, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f16xf16+f32acc", "data": "synthetic",
               "config": cfg,
               "e2e": {"value": round(triples2 / (ms_e2e / 1e3), 1), "unit": B.UNIT, "h2d_bytes_per_step": int(h2d / K),
                       "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 4)},
               "gpu_launches": int(round(launches_per_step * K)), "cuda_graph": graphed,
               "cuda_graph_note": graph_note, "clocks": clocks, "roofline": roof, "multi_gpu_parity": parity,
               "collectives_per_step": ["all_reduce X[B,D] f32", "all_reduce dQ[B,D] f32", "all_reduce loss f32"],
               "prefix_rows_per_sec": round(K * Bg / (ms_total / 1e3), 1)}
        return out
    return None


def run_data_parallel(args, rank, world, device, workload, wl, parity=None):
    """Batch-shared candidate lists (the OLPBench training configurations) over N GPUs through the public API: the same
    ``Models`` class under ``Trainer`` in a torch.distributed job. Every rank draws its own batches (``wl["batch"]`` prefix
    rows) and its own candidate list, runs the step on its replica and the gradients are averaged over the ranks (NCCL
    all-reduce inside the step's CUDA graph) before the optimizer step. Weak scaling: rows per GPU fixed."""
    import bench as B
    from open_knowledge_graph_embeddings_b200 import _capi
    from open_knowledge_graph_embeddings_b200 import dataset as D
    from open_knowledge_graph_embeddings_b200.trainer import Trainer

    wl, spec, model, train, valid = B.build_workload(workload, device, world, rank)        # same seed: identical replicas
    targs = {"optimization_config": {"optimizer": "Adagrad", "lr": wl["lr"], "weight_decay": wl["weight_decay"]},
             "lr_scheduler_config": None, "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True}
    trainer = Trainer(targs, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)
    assert trainer.data_parallel == (world > 1)
    trainer.model_with_loss.train()
    K, W, Bl = args.steps, args.warmup, wl["batch"]
    np.random.seed(1 + rank)                                   # the host collate samples negatives from numpy's global stream
    pool = B.make_batches(train, Bl, min(K + W, 12), seed=7 + 1000 * rank, pin=True)          # every rank its own batches
    dev_pool = [D.input_and_labels_to_device(b, True, device, non_blocking=False) for b in pool]
    timer = B.KernelTimer()
    if rank == 0:
        _capi.set_call_hook(timer.hook)
    sampler = B.ClockSampler(device.index)

    def eager(b, sync):
        for o in trainer.optimizers:
            o.update(trainer.epoch, trainer.training_steps)
        trainer.compute_one_batch(b, training=True, sync_loss=False)
        trainer.training_steps += 1
        if sync:
            trainer.last_loss.item()

    for i in range(W):
        eager(dev_pool[i % len(dev_pool)], False)
    torch.cuda.synchronize()
    K0 = min(K, 10)
    timer.enabled = rank == 0
    for i in range(K0):
        eager(dev_pool[(W + i) % len(dev_pool)], False)
    torch.cuda.synchronize()
    timer.enabled = False
    launches_per_step = timer.launches / K0 if rank == 0 else 0

    graph_note, gstep = "disabled (--no-cuda-graph)", None
    if not args.no_cuda_graph:
        ok = torch.ones(1, device=device)
        try:
            max_cand = (int(1.2 * max(int(b[6].numel()) for b in pool)) + 255) // 256 * 256
            gstep = trainer.make_graphed_step(dev_pool[0], max(4096, 4 * max(int(b[3].idx.numel()) for b in pool)),
                                              max_candidates=max_cand)
            if gstep is None:
                ok.zero_()
                graph_note = "configuration not capturable, eager launches"
        except Exception as ex:  # noqa: BLE001
            ok.zero_()
            graph_note = f"capture failed, eager launches: {type(ex).__name__}: {str(ex)[:120]}"
            if rank == 0:
                import traceback
                traceback.print_exc()
        torch.cuda.synchronize()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0:
            graph_note = "whole step (kernels + NCCL gradient all-reduces) replayed as one CUDA graph per rank"
        else:
            gstep = None
    run = (lambda b, s: gstep.step(b, s)) if gstep is not None else (lambda b, s: eager(b, bool(s)))

    def timed(n_steps, body):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        triples = body(n_steps)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1), 0.0], device=device, dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        u = torch.tensor([triples], device=device, dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)                     # triples of all ranks
        return float(u.item()), float(t[0].item())

    def resident(n):
        triples = 0.0
        for i in range(n):
            b = dev_pool[(W + i) % len(dev_pool)]
            run(b, False)
            triples += b[2] / 2.0
        return triples

    for i in range(W):
        run(dev_pool[i % len(dev_pool)], False)
    if rank == 0:
        sampler.start()
    triples, ms_total = timed(K, resident)

    # end to end through Trainer.train_epoch: rows shuffled and collated on the device inside the step's graph (every rank
    # its own shuffle), loss D2H per step; without a graph: host collate in a prefetching loader + H2D
    trainer.args["cuda_graph"] = gstep is not None
    if gstep is not None:
        trainer._graphed_step = gstep
    state = {"h2d": 0}

    def e2e(n):
        done, triples = 0, 0.0
        nnz0 = int(gstep.collate.nnz_total) if (gstep is not None and hasattr(gstep, "collate")) else 0
        while done < n:
            if gstep is not None:
                loader = train.get_row_loader(shuffle=True, seed=17 + n + 1000 * rank)
            else:
                loader = train.get_loader(shuffle=True, drop_last=True, seed=17 + n + 1000 * rank, prefetch=4)

            def limited(it=loader, left=n - done):
                for k, b in enumerate(it):
                    if k >= left:
                        return
                    if not isinstance(b, D.DeviceRows):
                        state["h2d"] += D.batch_h2d_bytes(b)
                        state["t"] = state.get("t", 0.0) + b[2] / 2.0
                    yield b
            class Sized:
                def __len__(self_inner):
                    return n - done
                def __iter__(self_inner):
                    return limited()
            before = trainer.training_steps
            res = trainer.train_epoch(Sized())
            done += trainer.training_steps - before
            assert res["loss"].count > 0
            if hasattr(loader, "close"):
                loader.close()
        if gstep is not None and hasattr(gstep, "collate"):
            triples = (int(gstep.collate.nnz_total) - nnz0) / 2.0
        else:
            triples = state.pop("t", 0.0)
        return triples

    e2e(max(W, 3))
    state["h2d"] = 0
    triples2, ms_e2e = timed(K, e2e)
    clocks = sampler.stop() if rank == 0 else None
    # replicas must still agree after all those steps: every rank applied the same averaged gradients (what remains is the
    # summation-order noise of the float atomics inside each rank's own backward; train_epoch re-broadcasts periodically)
    drift = torch.zeros(1, device=device)
    for prm in model.parameters():
        ref = prm.data.clone()
        dist.broadcast(ref, src=0)
        drift = torch.maximum(drift, (prm.data - ref).abs().max().reshape(1) / ref.abs().max().clamp_min(1e-30))
    dist.all_reduce(drift, op=dist.ReduceOp.MAX)
    if rank == 0:
        roof = B.roofline_of(timer.summary(), B.load_peaks(), {}, workload)
        if roof:
            for k, v in roof["breakdown"].items():
                v["ms_per_step"] = round(v["total_ms"] / K0, 4)
            roof["breakdown_note"] = f"CUDA events around every native call over {K0} eagerly launched steps on rank 0"
        cfg = B.config_of(workload, wl, world, wl["batch"])
        cfg["parallelism"] = f"data-parallel x{world}: every rank its own {wl['batch']}-row batch and candidate list, replicas of all tables"
        cfg["api"] = "Models.%s under Trainer (Trainer.data_parallel: gradient all-reduce before OptimRegime.step)" % wl["model"]
        grad_bytes = sum(p.numel() * 4 for p in model.parameters())
        exchange = f"all_reduce(avg) of every parameter gradient, {grad_bytes / 1e6:.0f} MB fp32 per rank"
        if getattr(trainer, "sparse_exchange", False):
            width = model.entity_embedding.weight.size(1)
            caps = sorted(getattr(gstep, "captured_capacities", [])) if gstep is not None else []
            rows = getattr(gstep, "last_union_rows", None) if gstep is not None else None
            exchange = (f"touched-row exchange of the token-table gradients: all_reduce(max) of {trainer.union_state()['rows_total']} "
                        f"touch flags, all_reduce(avg) of the union's rows (last step: {rows} rows; captured capacities {caps} rows "
                        f"= {[round(c * width * 4 / 1e6) for c in caps]} MB) instead of {grad_bytes / 1e6:.0f} MB dense; small "
                        f"parameters in one flat all_reduce")
        out = {"metric": B.METRIC, "value": round(triples / (ms_total / 1e3), 1), "unit": B.UNIT, "n_gpus": world, "steps": K,
               "warmup": W, "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f16xf16+f32acc", "data": "synthetic", "config": cfg,
               "e2e": {"value": round(triples2 / (ms_e2e / 1e3), 1), "unit": B.UNIT, "h2d_bytes_per_step": int(state["h2d"] / K),
                       "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 4),
                       "path": "Trainer.train_epoch over a per-rank shuffled row loader, collate on the device inside the graph"
                               if gstep is not None else "Trainer.train_epoch(get_loader(shuffle=True, prefetch=4))"},
               "gpu_launches": int(round(launches_per_step * K)), "cuda_graph": gstep is not None, "cuda_graph_note": graph_note,
               "clocks": clocks, "roofline": roof, "multi_gpu_parity": parity,
               "collectives_per_step": [exchange],
               "replica_max_relative_drift": float(drift.item()),
               "prefix_rows_per_sec": round(K * wl["batch"] * world / (ms_total / 1e3), 1)}
        return out
    return None


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
        return out
    return None
