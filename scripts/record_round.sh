#!/bin/bash
# Round-end record on one B200: bench line, reference arm, secondary benches, ncu launch list and
# two ncu --set full captures (forward step; backward kernels).  Outputs under gpurun_out/<tag>_*.
tag=${1:-rX}
REGEX='regex:nsf_tc_layer_kernel|kde1d_deposit_kernel|kde1d_finish|moments_kernel|kde1d_bwd_kernel|nsf_tc_dgrad_kernel|nsf_tc_wgrad_kernel'
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
echo "extras are part of the bench line (extra[])"
timeout 200 python scripts/kde2d_ab.py > gpurun_out/${tag}_kde2d_ab.txt 2>&1; echo "kde2d a/b rc=$?"
timeout 100 python scripts/prof_step.py > gpurun_out/${tag}_prof_step.log 2>&1; echo "prof_step rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-graph > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
# one launch of each kernel of the forward step (layer, moments, deposit, finish), then of the backward
# (finish_bwd, kde1d_bwd, layer<..,1>, dgrad, wgrad); the reports are reduced to their raw pages on the
# box (the .ncu-rep files with source exceed what travels back)
timeout 500 ncu --set full --clock-control none --import-source on -k "$REGEX" --launch-skip 12 -c 4 -o gpurun_out/${tag}_prof_fwd -f python scripts/prof_step.py > gpurun_out/${tag}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 700 ncu --set full --clock-control none --import-source on -k "$REGEX" --launch-skip 49 -c 5 -o gpurun_out/${tag}_prof_bwd -f python scripts/prof_step.py > gpurun_out/${tag}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
for w in fwd bwd; do
  ncu -i gpurun_out/${tag}_prof_$w.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_$w.csv 2> /dev/null
  rm -f gpurun_out/${tag}_prof_$w.ncu-rep
done
gzip -f gpurun_out/${tag}_launches.csv
du -sh gpurun_out
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['train_step']['ms_per_step'], d['cpu_baseline']['value'])
r=json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r.get('cpu_baseline'))
"
