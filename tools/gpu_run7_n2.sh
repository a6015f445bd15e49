#!/usr/bin/env bash
# 2 GPUs: multi-rank correctness tests, the bench line at N=2 (parity self-check inside), training schedule A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/r2_gpu_tests_n2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_gpu_tests_n2.log
run2() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run2 29611 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --train-mb-per-stage 4 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
run2 29612 tools/bench_train.py --gpus 2 --steps 3 --train-mb-per-stage 2 > gpurun_out/r2_train_n2_mb2.json 2> gpurun_out/r2_train_n2_mb2.err; echo "train mb2 rc=$?"
run2 29613 tools/bench_train.py --gpus 2 --steps 3 --train-mb-per-stage 8 > gpurun_out/r2_train_n2_mb8.json 2> gpurun_out/r2_train_n2_mb8.err; echo "train mb8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('bench n2', round(d['value'],1), d['unit'], 'parity', d.get('parity_check'))
print(' pipeline', d.get('pipeline'))
print(' train', round(d['train']['value'],2), d['train']['ms_per_step'], d['train']['roofline']['whole_step']['frac'])
for f in ('mb2','mb8'):
    t=json.loads(open(f'gpurun_out/r2_train_n2_{f}.json').read().strip().splitlines()[-1]); print(f, round(t['value'],2), 'samples/s', round(t['ms_per_step'],1), 'ms', t['roofline']['whole_step']['frac'])
PY
