#!/usr/bin/env bash
# 8 GPUs: training step with the split head backward — micro-batches per stage 4 / 2 / 8, and BASELINE config 4
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29713 tools/bench_train.py --gpus 8 --steps 3 --train-mb-per-stage 4 > gpurun_out/r2_train_n8s_mb4.json 2> gpurun_out/r2_train_n8s_mb4.err; echo "train mb4 rc=$?"
run 29714 tools/bench_train.py --gpus 8 --steps 3 --train-mb-per-stage 2 > gpurun_out/r2_train_n8s_mb2.json 2> gpurun_out/r2_train_n8s_mb2.err; echo "train mb2 rc=$?"
run 29715 tools/bench_train.py --gpus 8 --steps 3 --train-mb-per-stage 8 > gpurun_out/r2_train_n8s_mb8.json 2> gpurun_out/r2_train_n8s_mb8.err; echo "train mb8 rc=$?"
run 29716 tools/bench_train.py --gpus 8 --steps 3 --train-model Qwen/Qwen3-8B --train-batch 2 --train-seq 1024 --train-mb-per-stage 2 > gpurun_out/r2_train_cfg4s.json 2> gpurun_out/r2_train_cfg4s.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
for f in ('n8s_mb4','n8s_mb2','n8s_mb8','cfg4s'):
    try:
        t=json.loads(open(f'gpurun_out/r2_train_{f}.json').read().strip().splitlines()[-1]); print(f, round(t['value'],2), 'samples/s', round(t['ms_per_step'],1), 'ms', t['roofline']['whole_step']['frac'])
    except Exception as e: print(f, e)
PY
