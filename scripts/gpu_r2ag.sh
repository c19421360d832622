#!/bin/bash
# round 2, session AG: two compute warpgroups with 224 registers against three with 160
mkdir -p gpurun_out
MENTFLOW_B200_LIB=$PWD/variants/lib_wg2.so timeout 600 python -m pytest tests/test_gpu_nsf.py -m gpu -x -q 2>&1 | tail -3
bash scripts/ab_bench.sh default variants/lib_wg2.so default variants/lib_wg2.so 2>&1 | tee gpurun_out/r2ag_ab.txt
