#!/bin/bash
# Scaling runs of the three multi-GPU workloads on N GPUs of one box: ./scripts/run_scaling.sh N  (writes gpurun_out/scale_*_nN.json)
N=$1
for w in c3_lookup_distmult_1m c4_olpbench_unigram c5_olpbench_eval; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --workload $w > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  echo "$w rc=$?"; tail -c 400 gpurun_out/scale_${w}_n$N.err | tail -3
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${w}_n$N.json").read().strip().splitlines()[-1])
    print(d["config"].get("workload"), "N=", d["n_gpus"], d["metric"], d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
except Exception as e: print("no json", e)
PY
done
