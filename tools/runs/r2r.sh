#!/bin/bash
OUT=gpurun_out/${1:-r2r}; mkdir -p $OUT
timeout 200 python tools/oz_probe.py collector > $OUT/collector.jsonl 2>&1; cat $OUT/collector.jsonl
timeout 200 python tools/oz_probe.py timeline 2048 512 > $OUT/timeline.jsonl 2>&1; grep oz_timeline $OUT/timeline.jsonl
timeout 700 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python tools/gpu_probe.py potrf2 > $OUT/potrf2.jsonl 2>&1; cat $OUT/potrf2.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-900 $OUT/bench1.json; tail -3 $OUT/bench1.err
