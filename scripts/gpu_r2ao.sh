#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ao_ment.csv python scripts/prof_ment.py > gpurun_out/r2ao_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY' | tee gpurun_out/r2ao_ment.txt
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2ao_ment.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hdr=rows[h]; tot=0
for r in rows[h+1:]:
    if len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); v=float(d["Metric Value"].replace(",",""))/1e3; tot+=v
    print("%9.1f us  %s" % (v, d["Kernel Name"][:100]))
print("total %.1f us" % tot)
PY
