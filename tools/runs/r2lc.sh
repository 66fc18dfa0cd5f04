#!/bin/bash
OUT=gpurun_out/${1:-r2lc}; mkdir -p $OUT
timeout 300 python tools/gpu_probe.py potrf2 > $OUT/potrf2.jsonl 2>&1; cut -c1-330 $OUT/potrf2.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python tools/cell_batch_probe.py 1000 2000 > $OUT/cell_batch.jsonl 2>&1; cut -c1-330 $OUT/cell_batch.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-330 $OUT/bench1.json; tail -3 $OUT/bench1.err
