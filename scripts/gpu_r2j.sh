#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nsf.py -q -k "coupling" 2>&1 | tail -15 > gpurun_out/r2j_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "n2 rc=$?"
tail -15 gpurun_out/r2j_tests.txt; tail -12 gpurun_out/r2j_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cross_rank_sum'], d['shard_parity']); print(d['extra'][0]['value'], d['extra'][0].get('shard_parity'))"
