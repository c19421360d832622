#!/bin/bash
mkdir -p gpurun_out
scripts/ab_bench.sh default variants/lib_rcp8.so default variants/lib_rcp8.so > gpurun_out/r2q_ab.txt 2>&1
cat gpurun_out/r2q_ab.txt
MENTFLOW_B200_LIB=$PWD/variants/lib_rcp8.so python -m pytest tests/test_gpu_nsf.py -q 2>&1 | tail -2
