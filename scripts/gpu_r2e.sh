#!/bin/bash
mkdir -p gpurun_out
scripts/ab_bench.sh default variants/lib_poll500.so variants/lib_poll2000.so variants/lib_trywait.so default > gpurun_out/r2e_ab.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err
cat gpurun_out/r2e_ab.txt
