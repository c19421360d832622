#!/bin/bash
# Build a variant of the library with extra nvcc flags: scripts/build_variant.sh NAME "-DMFB_SPLINE_COMP=1 ..."
# -> variants/lib_NAME.so (git-ignored, travels to the GPU box); select it with MENTFLOW_B200_LIB.
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/variants; bd=$out/build_$name
mkdir -p "$bd"
pids=()
for src in "$root"/mentflow_b200/csrc/*.cu; do
  o=$bd/$(basename "${src%.cu}").o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 549 $flags -c "$src" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/lib_$name.so" "$bd"/*.o
rm -rf "$bd"
echo "$out/lib_$name.so"
