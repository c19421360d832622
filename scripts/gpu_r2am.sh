#!/bin/bash
mkdir -p gpurun_out
sed -n '/^cat > \/tmp\/k2b.py/,/^PY$/p' scripts/gpu_r2al.sh | sed '1d;$d' > /tmp/k2b.py
echo "previous:"; MENTFLOW_B200_LIB=$PWD/variants/lib_prev.so timeout 200 python /tmp/k2b.py 2>&1 | tail -1
echo "new:"; timeout 200 python /tmp/k2b.py 2>&1 | tail -1
