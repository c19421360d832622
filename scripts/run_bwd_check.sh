timeout 400 python -m pytest tests/test_gpu_nsf.py -x -q -k "backward" 2>&1 | tail -3; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nsf_tc" --csv --log-file gpurun_out/bwd_times.csv python scripts/bwd_prof.py > /dev/null 2>&1; python - <<PY
import csv,collections
lines=[l for l in open("gpurun_out/bwd_times.csv") if l.startswith("\"")]
r=csv.DictReader(lines); agg=collections.defaultdict(list)
for x in r: agg[x["Kernel Name"][:60]].append(float(x["Metric Value"].replace(",","")))
for k,v in agg.items(): print(k, len(v), "median us", sorted(v)[len(v)//2]/1e3)
PY
