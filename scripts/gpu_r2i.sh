#!/bin/bash
# 8 GPUs: weak scaling N = 8, 4 (flow headline, e2e, shard parity, MENT config 5 at 1e8 particles per update)
mkdir -p gpurun_out
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2i_bench_n$n.json 2> gpurun_out/r2i_bench_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f'gpurun_out/r2i_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d.get('shard_parity'))
        for e in d.get('extra', []):
            print('   ', e['case'][:60], '%.2f ms' % e['ms_per_step'], '%.4g' % e['value'], e.get('shard_parity'))
    except Exception as ex:
        print(n, 'ERR', ex)
PY
