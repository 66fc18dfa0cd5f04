#!/bin/bash
# ncu evidence of round 2 (one GPU): complete launch list of one fit+predict pass at N=40000, --set full captures of the
# dominant int8 kernel and of the HBM-bound kernels.  Outputs under gpurun_out/$1/.
OUT=gpurun_out/${1:-r2p}; mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_40k.csv python tools/one_fit.py 40000 > $OUT/launches_40k.log 2>&1
tail -2 $OUT/launches_40k.log
F="--set full --clock-control none --import-source on"
timeout 300 ncu $F -k regex:oz_mma -s 2 -c 1 -o $OUT/oz_mma_syrk16k_k2048 python tools/oz_probe.py one 8 2048 > $OUT/ncu_oz.log 2>&1
timeout 300 ncu $F -k regex:oz_mma -s 14 -c 1 -o $OUT/oz_mma_in_potrf python tools/one_fit.py 40000 > $OUT/ncu_oz_potrf.log 2>&1
timeout 300 ncu $F -k regex:cov_build -c 1 -o $OUT/cov_build_40k python tools/one_fit.py 40000 > $OUT/ncu_cov.log 2>&1
timeout 300 ncu $F -k regex:trsv_fwd_step -s 150 -c 1 -o $OUT/trsv_fwd_step_40k python tools/one_fit.py 40000 > $OUT/ncu_trsv.log 2>&1
timeout 300 ncu $F -k regex:trsv_bwd_step -s 150 -c 1 -o $OUT/trsv_bwd_step_40k python tools/one_fit.py 40000 > $OUT/ncu_trsvb.log 2>&1
timeout 300 ncu $F -k regex:oz_slice -s 8 -c 1 -o $OUT/oz_slice_40k python tools/one_fit.py 40000 > $OUT/ncu_slice.log 2>&1
ls -la $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "bit_identical" > $OUT/pytest_new.log 2>&1; tail -3 $OUT/pytest_new.log
