#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$" | tail -25 > gpurun_out/r2k_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "n2 rc=$?"
tail -12 gpurun_out/r2k_tests.txt; tail -4 gpurun_out/r2k_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cross_rank_sum'], d['shard_parity']); print(d['extra'][0]['value'], d['extra'][0].get('shard_parity'))"
