#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "n2 rc=$?"
python scripts/kde2d_ab.py > gpurun_out/r2g_kde2d_default.txt 2>&1
MENTFLOW_B200_LIB=$PWD/variants/lib_skipzero.so python scripts/kde2d_ab.py > gpurun_out/r2g_kde2d_skipzero.txt 2>&1
tail -4 gpurun_out/r2g_kde2d_default.txt; tail -4 gpurun_out/r2g_kde2d_skipzero.txt
python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['shard_parity']); print(d['extra'][0])"
