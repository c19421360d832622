#!/bin/bash
# round 2, session AH: dependent launch along the backward chain: tests + optimisation-step time at small batches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nsf.py tests/test_gpu_entropy_loss.py tests/test_gpu_kde1d.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2ah_tests.txt
for rep in 1 2; do
echo "== MFB_PDL=0"; MENTFLOW_B200_LIB=$PWD/variants/lib_nopdl.so timeout 300 python scripts/opt_step_time.py 25000 100000 1000000 2>&1 | tail -3
echo "== default";   timeout 300 python scripts/opt_step_time.py 25000 100000 1000000 2>&1 | tail -3
done | tee gpurun_out/r2ah_opt.txt
