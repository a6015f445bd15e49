#!/usr/bin/env bash
# 8 GPUs: the bench line exactly as the driver's scaling run launches it (parity self-check + generate + training step)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8b.json').read().strip().splitlines()[-1])
print('bench n8', round(d['value'],1), d['unit'], 'parity:', d.get('parity_check'))
print(' pipeline', d.get('pipeline'))
t=d['train']; print(' train', round(t['value'],2), round(t['ms_per_step'],1), t['roofline']['whole_step']['frac'])
PY
