"""Per-call device times and host time of the evaluation step (Trainer.compute_one_batch(training=False)) of a bench workload.

    python scripts/profile_eval_step.py [workload]     (default c3_lookup_distmult_1m; run on a B200 through gpurun)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from open_knowledge_graph_embeddings_b200 import _capi  # noqa: E402
from open_knowledge_graph_embeddings_b200.trainer import Trainer  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else B.DEFAULT_WORKLOAD
    device = torch.device("cuda", 0)
    wl, spec, model, train, valid = B.build_workload(workload, device, 1, 0)
    targs = {"optimization_config": {"optimizer": "Adagrad", "lr": wl["lr"], "weight_decay": wl["weight_decay"]},
             "lr_scheduler_config": None, "bce_label_smoothing": 0.0, "grad_clip": 0}
    trainer = Trainer(targs, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)
    trainer.model_with_loss.eval()
    batches = B.make_batches(valid, wl["batch"], 8, seed=11, pin=True)
    timer = B.KernelTimer()
    _capi.set_call_hook(timer.hook)
    with torch.no_grad():
        for b in batches[:3]:
            trainer.compute_one_batch(b, training=False)
        torch.cuda.synchronize()
        timer.enabled = True
        t0 = time.perf_counter()
        for b in batches[3:]:
            print("answers", len(b[4]), "overflow", int(b[4].overflow.numel()))
            trainer.compute_one_batch(b, training=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / len(batches[3:])
        timer.enabled = False
        # pipelined variant without per-call events
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        trainer.evaluate(batches)
        torch.cuda.synchronize()
        dt2 = (time.perf_counter() - t0) / len(batches)
    print(f"wall per step (synchronous, with per-call events): {dt * 1e3:.3f} ms; pipelined evaluate(): {dt2 * 1e3:.3f} ms")
    more = B.make_batches(valid, wl["batch"], 16, seed=12, pin=True)
    for rep in range(2):
        st0 = torch.cuda.memory_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            trainer.evaluate(more)
        torch.cuda.synchronize()
        st1 = torch.cuda.memory_stats()
        print(f"evaluate() over 16 new batches, pass {rep}: {(time.perf_counter() - t0) / 16 * 1e3:.3f} ms/step, "
              f"cudaMalloc calls {st1['num_device_alloc'] - st0['num_device_alloc']}, frees {st1['num_device_free'] - st0['num_device_free']}, "
              f"reserved {st1['reserved_bytes.all.current'] / 2**30:.1f} GiB")
    n = len(batches[3:])
    for k, v in sorted(timer.summary().items(), key=lambda kv: -kv[1]["ms"]):
        print(f"{v['ms'] / n:8.4f} ms/step  {v['calls'] / n:5.1f} calls  {k}")


if __name__ == "__main__":
    main()
