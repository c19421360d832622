#!/bin/bash
mkdir -p gpurun_out
MENTFLOW_B200_LIB=$PWD/variants/lib_reorder.so python -m pytest tests/test_gpu_nsf.py -q -x 2>&1 | tail -4 > gpurun_out/r2h_tests_reorder.txt
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2h_tests.txt
scripts/ab_bench.sh default variants/lib_reorder.so default variants/lib_reorder.so > gpurun_out/r2h_ab.txt 2>&1
cat gpurun_out/r2h_tests_reorder.txt gpurun_out/r2h_tests.txt gpurun_out/r2h_ab.txt
