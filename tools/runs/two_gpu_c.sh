#!/bin/bash
# 2-GPU run (gpurun --gpus 2): bench.py exactly as the driver launches it
OUT=gpurun_out/${1:-r2g2}; mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 800 $T bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_2gpu.json").read().strip().splitlines()[-1]); s = d["sharded"]
    print("per_gpu ms", d["ms_per_step"], "value", d["value"]); print("sharded ms", s.get("ms_per_step"), s.get("phase_seconds_rank0_serialised"), s.get("parity", {}).get("vs_cpu_oracle", {}).get("ok"), s.get("parity", {}).get("vs_single_gpu_engine", {}).get("ok"), s.get("error"))
except Exception as e:
    print("failed", e); print(open("$OUT/bench_2gpu.err").read()[-1500:])
PY
