#!/bin/bash
OUT=gpurun_out/${1:-r2s}; mkdir -p $OUT
./tools/microbench/l2_fill_peak > $OUT/l2_fill_peak.jsonl 2>&1; cat $OUT/l2_fill_peak.jsonl
timeout 200 python tools/oz_probe.py bound > $OUT/bound.jsonl 2>&1; cat $OUT/bound.jsonl
timeout 700 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_slice -s 8 -c 1 -o $OUT/oz_slice_big python tools/one_fit.py 40000 > $OUT/ncu_slice.log 2>&1; tail -2 $OUT/ncu_slice.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:oz_slice -c 40 --csv --log-file $OUT/slice_launches.csv python tools/one_fit.py 40000 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/slice_launches.csv")) if len(r)>10]
hdr=rows[0]; d={}
for r in rows[1:]:
    rec=dict(zip(hdr,r)); d.setdefault(rec["ID"],{"grid":rec["Grid Size"]})[rec["Metric Name"]]=float(rec["Metric Value"].replace(",",""))
for k,v in d.items():
    if v.get("gpu__time_duration.sum",0)>50000: print(k,v["grid"],v["gpu__time_duration.sum"],v["dram__bytes_read.sum"],v["dram__bytes_write.sum"])
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-700 $OUT/bench1.json; tail -3 $OUT/bench1.err
