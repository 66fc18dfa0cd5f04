#!/bin/bash
# 8-GPU run (gpurun --gpus 8): bench.py exactly as the driver launches it -- per-GPU batch (configs[3]) + sharded N=200k object (configs[4])
OUT=gpurun_out/${1:-r2g8}; mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 800 $T bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
tail -c 5000 $OUT/bench_8gpu.json; tail -5 $OUT/bench_8gpu.err
