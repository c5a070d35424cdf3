#!/bin/bash
# 8-GPU session: weak / strong scaling of the bench (config 2 and, under "also", config 5), eps-sweep batching under strong
# scaling, and the data-parallel training steps with the NCCL gradient all-reduce.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_weak.json 2> gpurun_out/bench_${N}gpu_weak.err; echo "weak rc=$?"
$TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/bench_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu_strong.err; echo "strong rc=$?"
$TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --eps-sweep 6 --no-config5 > gpurun_out/bench_${N}gpu_strong_sweep6.json 2> gpurun_out/bench_${N}gpu_strong_sweep6.err; echo "strong sweep rc=$?"
$TR tools/train_step_multi.py > gpurun_out/train_step_${N}gpu.json 2> gpurun_out/train_step_${N}gpu.err; echo "train rc=$?"
for f in weak strong strong_sweep6; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu_$f.json").read().strip().splitlines()[-1])
print("$f", d["n_gpus"], d["scaling"], round(d["value"]/1e6,1), "M pairs/s", round(d["ms_per_step"],2), "ms", {k: (round(v["value"]/1e6,1), round(v["ms_per_step"],1)) for k,v in d.get("also",{}).items()})
PY
done
cat gpurun_out/train_step_${N}gpu.json
