#!/bin/bash
OUT=gpurun_out/${1:-r2last}; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-400 $OUT/bench1.json; tail -2 $OUT/bench1.err
