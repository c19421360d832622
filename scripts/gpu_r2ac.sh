#!/bin/bash
# round 2, session AC: banded operand rows in the KDE-2D tensor-core kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kde2d.py tests/test_gpu_baseline_sized.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2ac_tests.txt
echo "== dense"; MENTFLOW_B200_LIB=$PWD/variants/lib_dense.so timeout 200 python scripts/kde2d_ab.py 2>&1 | tail -12 | tee gpurun_out/r2ac_dense.txt
echo "== band"; timeout 200 python scripts/kde2d_ab.py 2>&1 | tail -12 | tee gpurun_out/r2ac_band.txt
