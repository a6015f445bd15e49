#!/usr/bin/env bash
# usage: gpu_run_scale.sh N   — N-GPU validation + measurements (N = 4 or 8)
N=$1
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "$N" = "4" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -q -x -k "peer_ring" > gpurun_out/r2_gpu_tests_n4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_n4.log
  run 29711 bench.py --gpus 4 --workload cfg3 --steps 2 --warmup 3 --no-train --no-parity-check > gpurun_out/r2_bench_cfg3_n4.json 2> gpurun_out/r2_bench_cfg3_n4.err; echo "cfg3 n4 rc=$?"
fi
run 29712 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --train-mb-per-stage 2 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
run 29713 tools/bench_train.py --gpus $N --steps 3 --train-mb-per-stage 4 > gpurun_out/r2_train_n${N}_mb4.json 2> gpurun_out/r2_train_n${N}_mb4.err; echo "train mb4 rc=$?"
if [ "$N" = "8" ]; then
  run 29714 tools/bench_train.py --gpus 8 --steps 3 --train-model Qwen/Qwen3-8B --train-batch 2 --train-seq 1024 --train-mb-per-stage 2 > gpurun_out/r2_train_cfg4.json 2> gpurun_out/r2_train_cfg4.err; echo "cfg4 rc=$?"
  run 29715 tools/bench_train.py --gpus 8 --steps 3 --train-model Qwen/Qwen3-8B --train-batch 2 --train-seq 1024 --train-mb-per-stage 4 > gpurun_out/r2_train_cfg4_mb4.json 2> gpurun_out/r2_train_cfg4_mb4.err; echo "cfg4 mb4 rc=$?"
  run 29716 bench.py --gpus 8 --workload cfg5 --rows-per-gpu 4 --steps 2 --warmup 3 --no-train --no-parity-check --no-cpu-baseline > gpurun_out/r2_bench_cfg5_n8.json 2> gpurun_out/r2_bench_cfg5_n8.err; echo "cfg5 n8 rc=$?"
fi
python - <<PY
import json,glob
def last(f):
    try: return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: return None
d=last('gpurun_out/r2_bench_n$N.json')
if d:
    print('bench n$N', round(d['value'],1), d['unit'], 'parity ok:', d.get('parity_check',{}).get('ok'), d.get('parity_check'))
    print(' pipeline', d.get('pipeline'))
    t=d.get('train'); print(' train mb2', round(t['value'],2), round(t['ms_per_step'],1), t['roofline']['whole_step']['frac'])
for f in sorted(glob.glob('gpurun_out/r2_train_*n${N}*.json')+glob.glob('gpurun_out/r2_train_cfg4*.json')+glob.glob('gpurun_out/r2_bench_cfg*_n$N.json')):
    t=last(f)
    if t: print(f.split('/')[-1], round(t['value'],2), t['unit'], round(t['ms_per_step'],1), 'ms', (t.get('roofline') or {}).get('whole_step',{}).get('frac'), (t.get('pipeline') or {}).get('exposed_wait_frac_worst_rank'))
PY
