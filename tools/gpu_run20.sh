#!/usr/bin/env bash
# 1 GPU: tcgen05 attention backward timing (alone and inside the training step), then the training / parity tests that use it
mkdir -p gpurun_out
timeout 300 python tools/bench_attn_bwd.py > gpurun_out/r2_attn_bwd_bench.txt 2>&1; echo "bench attn_bwd rc=$?"; cat gpurun_out/r2_attn_bwd_bench.txt
TL_ATTN_BWD=mma timeout 400 python tools/bench_train.py --steps 3 > gpurun_out/r2_train_n1_bwd_mma.json 2> gpurun_out/r2_train_n1_bwd_mma.err; echo "train mma rc=$?"
timeout 400 python tools/bench_train.py --steps 3 > gpurun_out/r2_train_n1_bwd_tc.json 2> gpurun_out/r2_train_n1_bwd_tc.err; echo "train tc rc=$?"
for f in bwd_mma bwd_tc; do python -c "
import json
t=json.loads(open('gpurun_out/r2_train_n1_$f.json').read().strip().splitlines()[-1]); print('$f', round(t['value'],2), 'samples/s', round(t['ms_per_step'],1), 'ms', t['roofline']['whole_step']['frac'])" 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_scale_gpu.py -q -x -k "training or train or attn_bwd or upstream or deferred" > gpurun_out/r2_gpu_tests20.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests20.log
