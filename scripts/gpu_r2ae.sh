#!/bin/bash
# round 2, session AE: A hi operand of the flow kernel in tensor memory
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nsf.py tests/test_gpu_baseline_sized.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2ae_tests.txt
bash scripts/ab_bench.sh variants/lib_nots.so default variants/lib_nots.so default 2>&1 | tee gpurun_out/r2ae_ab.txt
