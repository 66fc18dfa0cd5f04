#!/bin/bash
OUT=gpurun_out/${1:-r2lf}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_exactgp.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 300 python tools/gpu_probe.py chaincfg > $OUT/chaincfg.jsonl 2>&1; cut -c1-200 $OUT/chaincfg.jsonl
cp battgp_b200/lib/libbattgp_b200_prof.so battgp_b200/lib/libbattgp_b200.so
timeout 200 python tools/gpu_probe.py leaf > $OUT/leaf.jsonl 2>&1; cut -c1-330 $OUT/leaf.jsonl | head -3
