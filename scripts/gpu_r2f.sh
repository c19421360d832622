#!/bin/bash
# two GPUs: packed all-reduce, shard parity (flow + MENT), N=2 bench line; new BASELINE-sized parity tests
mkdir -p gpurun_out
python -m pytest tests/test_gpu_baseline_sized.py -q 2>&1 | grep -v "^  \|^$" | tail -30 > gpurun_out/r2f_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "n2 rc=$?"
scripts/ab_bench.sh default variants/lib_poll0.so > gpurun_out/r2f_ab.txt 2>&1
tail -20 gpurun_out/r2f_tests.txt; tail -5 gpurun_out/r2f_bench_n2.err; cat gpurun_out/r2f_ab.txt
