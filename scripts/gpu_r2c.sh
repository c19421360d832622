#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$" | tail -60 > gpurun_out/r2c_tests.txt
scripts/ab_bench.sh default variants/lib_split1.so variants/lib_split2.so variants/lib_split3.so variants/lib_comp0.so default > gpurun_out/r2c_ab.txt 2>&1
for v in split3; do
  MENTFLOW_B200_LIB=$PWD/variants/lib_$v.so python -m pytest tests/test_gpu_nsf.py -q 2>&1 | tail -3 > gpurun_out/r2c_tests_$v.txt
done
tail -25 gpurun_out/r2c_tests.txt; cat gpurun_out/r2c_ab.txt; cat gpurun_out/r2c_tests_split3.txt
