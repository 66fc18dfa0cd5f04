#!/bin/bash
OUT=gpurun_out/${1:-r2t}; mkdir -p $OUT
./tools/microbench/i8_peak 0.5 > $OUT/i8_peak.jsonl 2>&1; cat $OUT/i8_peak.jsonl
timeout 200 python tools/oz_probe.py bound > $OUT/bound.jsonl 2>&1; cat $OUT/bound.jsonl
timeout 700 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:oz_slice -c 20 --csv --log-file $OUT/slice_launches.csv python tools/one_fit.py 40000 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/slice_launches.csv")) if len(r)>10]
hdr=rows[0]; d={}
for r in rows[1:]:
    rec=dict(zip(hdr,r)); d.setdefault(rec["ID"],{"grid":rec["Grid Size"]})[rec["Metric Name"]]=float(rec["Metric Value"].replace(",",""))
for k,v in d.items():
    if v.get("gpu__time_duration.sum",0)>30000: print(k,v["grid"],v)
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1.json 2> $OUT/bench1.err; cut -c1-700 $OUT/bench1.json; tail -3 $OUT/bench1.err
