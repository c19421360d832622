#!/bin/bash
# A/B of library variants: scripts/ab_bench.sh variants/lib_a.so variants/lib_b.so ...  ("default" = in-tree lib)
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset MENTFLOW_B200_LIB; else export MENTFLOW_B200_LIB=$PWD/$lib; fi
  timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$lib" <<PY
import json,sys
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4g step %.4f ms  nsf %.4f ms  rest %.4f  e2e %.4g  train %.3f ms" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["entropy_project_kde_loss_ms_per_step"], d["e2e"]["value"], d["train_step"]["ms_per_step"]))
PY
done
