#!/bin/bash
# 8 GPUs with the NVLink peer exchange: weak scaling N = 8, 4 (+ the coupling parity test on one GPU)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nsf.py -q -k coupling 2>&1 | tail -6 > gpurun_out/r2m_tests.txt
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2m_bench_n$n.json 2> gpurun_out/r2m_bench_n$n.err; echo "n$n rc=$?"
done
cat gpurun_out/r2m_tests.txt
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f'gpurun_out/r2m_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['config']['cross_rank_sum'][:30], d.get('shard_parity', {}).get('ok'), d.get('shard_parity', {}).get('max_rel_profile'))
        for e in d.get('extra', []):
            print('   ', e['case'][:40], '%.2f ms' % e['ms_per_step'], '%.4g' % e['value'], e.get('shard_parity', {}).get('max_rel_profile'))
    except Exception as ex:
        print(n, 'ERR', ex)
PY
tail -3 gpurun_out/r2m_bench_n8.err
