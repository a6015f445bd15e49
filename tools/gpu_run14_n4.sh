#!/usr/bin/env bash
# 4 GPUs: the bench line as the driver's scaling run launches it (generate + training with the split head + parity self-check)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n4b.json 2> gpurun_out/r2_bench_n4b.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4b.json').read().strip().splitlines()[-1])
print('bench n4', round(d['value'],1), d['unit'], 'parity ok:', d.get('parity_check',{}).get('ok'))
t=d['train']; print(' train', round(t['value'],2), round(t['ms_per_step'],1), t['roofline']['whole_step']['frac'], t['config']['workload'][:120])
PY
