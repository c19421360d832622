#!/bin/bash
# round 2, session Z2: dgrad next-tile prefetch + wgrad staging depth 3: tests, kernel times, training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nsf.py tests/test_gpu_entropy_loss.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2z2_tests.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nsf_tc_(layer|dgrad|wgrad)_kernel' --launch-skip 15 -c 15 --csv --log-file gpurun_out/r2z2.csv python scripts/bwd_prof.py > gpurun_out/r2z2_ncu.log 2>&1
python - <<'PY' | tee gpurun_out/r2z2_kernels.txt
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2z2.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hdr=rows[h]; agg=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); agg[d["Kernel Name"].split("(")[0][-34:]].append(float(d["Metric Value"].replace(",","")))
for k,v in agg.items(): print(k, round(sum(v)/len(v)/1e6,4), "ms", len(v))
PY
bash scripts/ab_bench.sh variants/lib_prev.so default variants/lib_prev.so default 2>&1 | tee gpurun_out/r2z2_ab.txt
