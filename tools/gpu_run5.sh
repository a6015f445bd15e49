#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/trace_chain.py > gpurun_out/r2_trace_v4.txt 2>&1; echo "trace rc=$?"
tail -31 gpurun_out/r2_trace_v4.txt
B="--no-train --no-cpu-baseline --no-parity-check --steps 2 --warmup 3"
timeout 300 python bench.py $B > gpurun_out/r2_ab_v4.json 2> gpurun_out/r2_ab_v4.err; echo "bench rc=$?"
TL_CHAIN_STAGE_KB=16 timeout 300 python bench.py $B > gpurun_out/r2_ab_v4_s16.json 2> gpurun_out/r2_ab_v4_s16.err; echo "bench16 rc=$?"
for f in ab_v4 ab_v4_s16; do python -c "
import json
d=json.loads(open('gpurun_out/r2_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), d['unit'])" 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_decode_chain_gpu.py tests/test_sampling_gpu.py tests/test_api_gpu.py tests/test_worker_gpu.py -q > gpurun_out/r2_gpu_tests5.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_gpu_tests5.log
