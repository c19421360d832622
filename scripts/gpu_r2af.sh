#!/bin/bash
# round 2, final code: weak scaling on 8 GPUs (N = 8, 4, 2) with the NVLink peer exchange + reference arm
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29650 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2af_bench_n$n.json 2> gpurun_out/r2af_bench_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
for n in (8, 4, 2):
    try:
        d = json.loads(open(f'gpurun_out/r2af_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['config']['cross_rank_sum'][:30], d.get('shard_parity'))
        for e in d.get('extra', []):
            print('   ', e['case'][:40], '%.2f ms' % e['ms_per_step'], '%.4g' % e['value'], e.get('shard_parity', {}))
    except Exception as ex:
        print(n, 'ERR', ex)
PY
tail -3 gpurun_out/r2af_bench_n8.err
