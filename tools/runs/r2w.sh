#!/bin/bash
OUT=gpurun_out/${1:-r2w}; mkdir -p $OUT
timeout 200 python tools/oz_probe.py bound > $OUT/bound.jsonl 2>&1; cat $OUT/bound.jsonl
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-700 $OUT/bench1.json; tail -3 $OUT/bench1.err
