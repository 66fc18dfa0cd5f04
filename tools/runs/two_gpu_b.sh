#!/bin/bash
# 2-GPU validation of the augmented-row sharded path: NCCL tests + sharded bench at N=$2 with phases
OUT=gpurun_out/${1:-r2f}; mkdir -p $OUT
SN=${2:-80000}
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded.log 2>&1; tail -3 $OUT/pytest_sharded.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 2 --workload sharded --size $SN --nb 1024 --steps 2 --warmup 1 --phases --verify > $OUT/sharded.json 2> $OUT/sharded.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/sharded.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["detail"].get("phase_seconds_rank0_serialised"), d["detail"]["residual_Kalpha_minus_y_over_y"], d["detail"]["parity"])
except Exception as e:
    print("failed", e); print(open("$OUT/sharded.err").read()[-1500:])
PY
