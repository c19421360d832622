#!/bin/bash
mkdir -p gpurun_out
for v in norot rot norot rot; do
  export MENTFLOW_B200_LIB=$PWD/variants/lib_$v.so
  python scripts/kde2d_ab.py 2>&1 | grep "tensor cores" | sed "s/^/$v: /"
done
MENTFLOW_B200_LIB=$PWD/variants/lib_rot.so python -m pytest tests/test_gpu_kde2d.py tests/test_gpu_baseline_sized.py -q 2>&1 | tail -1
