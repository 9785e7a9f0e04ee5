"""Host-side profile (cProfile) of Trainer.compute_one_batch for a small, launch-bound workload."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from open_knowledge_graph_embeddings_b200 import dataset as D  # noqa: E402
from open_knowledge_graph_embeddings_b200.trainer import Trainer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c1_fb15k237_complex"
device = torch.device("cuda")
wl, spec, model, train, valid = bench.build_workload(workload, device, 1, 0)
targs = {"optimization_config": {"optimizer": "Adagrad", "lr": wl["lr"], "weight_decay": wl["weight_decay"]},
         "lr_scheduler_config": None, "bce_label_smoothing": 0.0, "grad_clip": 0, "fused_entity_update": True}
trainer = Trainer(targs, model, torch.nn.BCEWithLogitsLoss(reduction="sum"), train, valid)
trainer.model_with_loss.train()
pool = [D.input_and_labels_to_device(b, True, device, non_blocking=False) for b in bench.make_batches(train, wl["batch"], 8, 7, True)]


def step(b):
    for o in trainer.optimizers:
        o.update(trainer.epoch, trainer.training_steps)
    trainer.compute_one_batch(b, training=True, sync_loss=False)
    trainer.training_steps += 1


for i in range(5):
    step(pool[i % 8])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(50):
    step(pool[i % 8])
t_launch = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
print(f"{workload}: host launch time {t_launch / 50 * 1e3:.3f} ms/step, with GPU drain {t_total / 50 * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(50):
    step(pool[i % 8])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

if len(sys.argv) > 2 and sys.argv[2] == "eval":
    trainer.model_with_loss.eval()
    ev = [D.input_and_labels_to_device(b, False, device, non_blocking=False) for b in bench.make_batches(valid, wl["batch"], 8, 11, True)]
    with torch.no_grad():
        for i in range(3):
            trainer.compute_one_batch(ev[i], training=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(20):
            trainer.compute_one_batch(ev[i % 8], training=False)
        torch.cuda.synchronize()
        print(f"eval: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/batch")
        pr = cProfile.Profile()
        pr.enable()
        for i in range(20):
            trainer.compute_one_batch(ev[i % 8], training=False)
        pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
