#!/bin/bash
mkdir -p gpurun_out
for v in new new2 new new2; do
  export MENTFLOW_B200_LIB=$PWD/variants/lib_$v.so
  python scripts/kde2d_ab.py 2>&1 | grep "tensor cores\|bins in (0.001\|bins in (1e-06\|bins in (1e-09" | sed "s/^/$v: /"
done
MENTFLOW_B200_LIB=$PWD/variants/lib_new2.so python -m pytest tests/test_gpu_kde2d.py tests/test_gpu_baseline_sized.py -q 2>&1 | tail -3
