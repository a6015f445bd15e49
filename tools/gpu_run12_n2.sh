#!/usr/bin/env bash
# 2 GPUs: split-head training step — single-GPU training tests, multi-rank tests, bench at N=2, micro-batch A/B, Adam overlap A/B
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_train_gpu.py -q -x -m gpu > gpurun_out/r2_gpu_tests_train12.log 2>&1; echo "pytest train rc=$?"
tail -5 gpurun_out/r2_gpu_tests_train12.log
timeout 1200 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/r2_gpu_tests_n2b.log 2>&1; echo "pytest n2 rc=$?"
tail -5 gpurun_out/r2_gpu_tests_n2b.log
run2() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run2 29611 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench n2 rc=$?"
run2 29612 tools/bench_train.py --gpus 2 --steps 3 --train-mb-per-stage 4 > gpurun_out/r2_train_n2b_mb4.json 2> gpurun_out/r2_train_n2b_mb4.err; echo "train mb4 rc=$?"
CUDA_VISIBLE_DEVICES=0 TL_ADAM_OVERLAP=0 timeout 400 python tools/bench_train.py --steps 3 > gpurun_out/r2_train_n1_serial.json 2> gpurun_out/r2_train_n1_serial.err; echo "train n1 serial rc=$?"
CUDA_VISIBLE_DEVICES=0 TL_ADAM_OVERLAP=1 timeout 400 python tools/bench_train.py --steps 3 > gpurun_out/r2_train_n1_overlap.json 2> gpurun_out/r2_train_n1_overlap.err; echo "train n1 overlap rc=$?"
python - <<'PY'
import json
def last(f): return json.loads(open(f).read().strip().splitlines()[-1])
try:
    d=last('gpurun_out/r2_bench_n2b.json')
    print('bench n2', round(d['value'],1), d['unit'], 'parity', d.get('parity_check'))
    print(' train', round(d['train']['value'],2), d['train']['ms_per_step'], d['train']['roofline']['whole_step']['frac'])
except Exception as e: print('bench n2', e)
for f in ('n2b_mb4','n1_serial','n1_overlap'):
    try:
        t=last(f'gpurun_out/r2_train_{f}.json'); print(f, round(t['value'],2), 'samples/s', round(t['ms_per_step'],1), 'ms', t['roofline']['whole_step']['frac'])
    except Exception as e: print(f, e)
PY
