#!/bin/bash
OUT=gpurun_out/${1:-r2q}; mkdir -p $OUT
timeout 700 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
timeout 200 python tools/oz_probe.py l2hint > $OUT/l2hint.jsonl 2>&1; cat $OUT/l2hint.jsonl
for h in 0 1 2 3; do timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:oz_mma -s 2 -c 1 --csv --log-file $OUT/ncu_l2hint$h.csv python tools/oz_probe.py one 8 2048 $h > $OUT/one_$h.log 2>&1; grep -E "dram__bytes_read|hit_rate|tensor|duration" $OUT/ncu_l2hint$h.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'; done
timeout 300 python tools/cell_batch_probe.py 1000 2000 4000 > $OUT/cell_batch.jsonl 2>&1; cat $OUT/cell_batch.jsonl
timeout 500 python tools/gpu_probe.py sched2 > $OUT/sched2.jsonl 2>&1; cat $OUT/sched2.jsonl
timeout 500 python tools/run_reference_modules.py --out $OUT/reference_modules.jsonl > $OUT/refmod.log 2>&1; tail -1 $OUT/reference_modules.jsonl | cut -c1-600
timeout 600 python bench.py --workload train --steps 3 --warmup 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; cut -c1-2500 $OUT/bench_train.json; tail -3 $OUT/bench_train.err
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-3500 $OUT/bench1.json; tail -3 $OUT/bench1.err
