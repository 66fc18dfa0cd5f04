#!/bin/bash
OUT=gpurun_out/${1:-r2cc}; mkdir -p $OUT
timeout 300 python tools/gpu_probe.py chaincfg > $OUT/chaincfg.jsonl 2>&1; cut -c1-300 $OUT/chaincfg.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 300 python tools/cell_batch_probe.py 1000 4000 > $OUT/cell_batch.jsonl 2>&1; cut -c1-330 $OUT/cell_batch.jsonl
