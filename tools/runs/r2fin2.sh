#!/bin/bash
# last validation of the round with the final library: smoke, GPU suite, size table, bench lines
OUT=gpurun_out/${1:-r2fin2}; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 300 python tools/gpu_probe.py potrf2 > $OUT/potrf2.jsonl 2>&1; cat $OUT/potrf2.jsonl
timeout 300 python tools/oz_probe.py raster > $OUT/raster.jsonl 2>&1; grep '"cluster": 1' $OUT/raster.jsonl
timeout 600 python bench.py > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-500 $OUT/bench1.json; tail -2 $OUT/bench1.err
python - <<PY
import json
d=json.loads(open("$OUT/bench1.json").read().strip().splitlines()[-1]); r=d["roofline"]
print({k:r[k] for k in ("achieved","peak","frac","traffic","kernel_fp64_equivalent_tflops","kernel_share_of_step")}, r["potrf"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
