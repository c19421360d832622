#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$" | tail -30 > gpurun_out/r2n_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -12 gpurun_out/r2n_tests.txt; tail -3 gpurun_out/r2n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['train_step'])
for e in d['extra']: print(e['case'][:70], e.get('particles_per_step'), round(e['ms_per_step'],3), '%.3g'%e['value'])"
