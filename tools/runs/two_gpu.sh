#!/bin/bash
# 2-GPU validation run (gpurun --gpus 2): NCCL parity tests, bench.py --gpus 2 exactly as the driver launches it (per-GPU
# batch + sharded N=200k object), look-ahead on/off at N=$2
OUT=gpurun_out/${1:-r2e}; mkdir -p $OUT
SN=${2:-80000}
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded.log 2>&1; tail -3 $OUT/pytest_sharded.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 800 $T bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -c 7000 $OUT/bench_2gpu.json; tail -5 $OUT/bench_2gpu.err
BATTGP_SHARDED_LOOKAHEAD=0 timeout 300 $T bench.py --gpus 2 --workload sharded --size $SN --nb 1024 --steps 2 --warmup 1 > $OUT/sharded_nolook.json 2> $OUT/sharded_nolook.err
timeout 300 $T bench.py --gpus 2 --workload sharded --size $SN --nb 1024 --steps 2 --warmup 1 --phases > $OUT/sharded_look.json 2> $OUT/sharded_look.err
python - <<PY
import json
for f in ("sharded_nolook", "sharded_look"):
    try:
        d = json.loads(open("$OUT/" + f + ".json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["detail"].get("phase_seconds_rank0_serialised"))
    except Exception as e:
        print(f, "failed", e); print(open("$OUT/" + f + ".err").read()[-1500:])
PY
