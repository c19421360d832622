#!/bin/bash
# round 2, session AA: launch list of one optimisation step at the reference batch size (25,000 particles)
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aa_small.csv python scripts/train_small_prof.py 25000 > gpurun_out/r2aa_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY' | tee gpurun_out/r2aa_small.txt
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2aa_small.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hdr=rows[h]; agg=collections.OrderedDict(); tot=0; cnt=0
for r in rows[h+1:]:
    if len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); k=d["Kernel Name"].split("(")[0][-48:]; v=float(d["Metric Value"].replace(",",""))/1e3
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v; cnt+=1
print("launches", cnt, "sum of kernel durations %.1f us" % tot)
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]: print("%8.1f us %4d  %s" % (v,c,k))
PY
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for e in d.get('extra',[]): print(e.get('workload'), e.get('ms_per_step'), e.get('value'))
" | tee gpurun_out/r2aa_extras.txt
