"""Host-side profile (cProfile) of the end-to-end training path of bench.py's `e2e` leg:
Trainer.train_epoch(dataset.get_loader(shuffle=True, prefetch=4)) with the step replayed as a CUDA graph."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from open_knowledge_graph_embeddings_b200.trainer import Trainer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c1_fb15k237_complex"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
device = torch.device("cuda")
wl, spec, model, train, valid = bench.build_workload(workload, device, 1, 0)
targs = {"optimization_config": {"optimizer": "Adagrad", "lr": wl["lr"], "weight_decay": wl["weight_decay"]},
         "lr_scheduler_config": None, "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True, "cuda_graph": True}
trainer = Trainer(targs, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)


def run(n, seed):
    loader = train.get_row_loader(shuffle=True, seed=seed) if wl.get("shared") else \
        train.get_loader(shuffle=True, drop_last=True, seed=seed, prefetch=4)
    t0 = time.perf_counter()
    trainer.train_epoch(loader, max_steps=n)
    torch.cuda.synchronize()
    if hasattr(loader, "close"):
        loader.close()
    return time.perf_counter() - t0


run(50, 1)
t = run(steps, 2)
print(f"{workload}: {t / steps * 1e3:.4f} ms/step end to end over {steps} steps")
pr = cProfile.Profile()
pr.enable()
run(steps, 3)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
