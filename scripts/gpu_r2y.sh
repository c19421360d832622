#!/bin/bash
# round 2, session Y: launch list + DRAM bytes of the backward kernels after the compact workspace
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'nsf_tc_(layer|dgrad|wgrad)_kernel' --launch-skip 15 -c 15 --csv --log-file gpurun_out/r2y_bwd.csv python scripts/bwd_prof.py > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2y_bwd.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hdr=rows[h]; agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[h+1:]:
    if len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); name=d["Kernel Name"].split("(")[0][-40:]
    agg[name][d["Metric Name"]].append(float(d["Metric Value"].replace(",","")))
for k,v in agg.items():
    print(k, {m: round(sum(x)/len(x),1) for m,x in v.items()}, "n=%d"%len(next(iter(v.values()))))
PY
