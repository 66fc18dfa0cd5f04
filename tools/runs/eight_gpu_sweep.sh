#!/bin/bash
# 8-GPU sweep of the sharded factorisation's overlap knobs at N=200k
OUT=gpurun_out/${1:-r2g8b}; mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $T bench.py --gpus 8 --workload sharded --size 200000 --nb 2048 --steps 2 --warmup 1 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1]); print("$name", round(d["ms_per_step"],1), "ms")
except Exception as e:
    print("$name failed", e); print(open("$OUT/$name.err").read()[-800:])
PY
}
run nolook BATTGP_SHARDED_LOOKAHEAD=0
run look_tpc4 BATTGP_SHARDED_TPC=4
run look_tpc16 BATTGP_SHARDED_TPC=16
run look_tpc64 BATTGP_SHARDED_TPC=64 BATTGP_SHARDED_TPC_SHORT=4
