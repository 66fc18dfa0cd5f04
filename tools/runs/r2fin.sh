#!/bin/bash
# final single-GPU validation of the round: smoke, full GPU test suite, the driver's bench lines
OUT=gpurun_out/${1:-r2fin}; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-600 $OUT/bench1.json; tail -2 $OUT/bench1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-400 $OUT/bench_ref.json
timeout 600 python bench.py --workload train --steps 3 --warmup 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; cut -c1-500 $OUT/bench_train.json
timeout 300 python tools/cell_batch_probe.py 1000 > $OUT/cell_batch.jsonl 2>&1; cat $OUT/cell_batch.jsonl
timeout 300 python tools/run_reference_modules.py --out $OUT/reference_modules.jsonl > $OUT/refmod.log 2>&1; tail -1 $OUT/reference_modules.jsonl | cut -c1-300
