#!/usr/bin/env bash
mkdir -p gpurun_out
B="--no-train --no-cpu-baseline --no-parity-check --steps 2 --warmup 3"
timeout 300 python bench.py $B > gpurun_out/r2_ab_k2.json 2> gpurun_out/r2_ab_k2.err; echo "kernels rc=$?"
TL_DECODE_IMPL=dq timeout 300 python bench.py $B > gpurun_out/r2_ab_dq.json 2> gpurun_out/r2_ab_dq.err; echo "dq rc=$?"
TL_DECODE_IMPL=dq TL_CHAIN_STAGE_KB=16 timeout 300 python bench.py $B > gpurun_out/r2_ab_dq16.json 2> gpurun_out/r2_ab_dq16.err; echo "dq16 rc=$?"
for f in ab_k2 ab_dq ab_dq16; do python -c "
import json
d=json.loads(open('gpurun_out/r2_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), d['unit'], d['gpu_launches'])" 2>&1 | tail -1; done
TL_DECODE_IMPL=dq timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_scale_gpu.py -q -x > gpurun_out/r2_gpu_tests8.log 2>&1; echo "pytest dq rc=$?"
tail -5 gpurun_out/r2_gpu_tests8.log
