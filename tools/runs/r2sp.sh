#!/bin/bash
OUT=gpurun_out/${1:-r2sp}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_exactgp.py tests/test_gpu_sharded.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python tools/gpu_probe.py chainsplit > $OUT/chainsplit.jsonl 2>&1; cut -c1-220 $OUT/chainsplit.jsonl
