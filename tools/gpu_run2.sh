#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/trace_chain.py > gpurun_out/r2_trace_chain.txt 2>&1; echo "trace rc=$?"
B="--no-train --no-cpu-baseline --no-parity-check --steps 2 --warmup 3"
TL_DECODE_IMPL=kernels timeout 300 python bench.py $B > gpurun_out/r2_ab_kernels.json 2> gpurun_out/r2_ab_kernels.err; echo "kernels rc=$?"
TL_DECODE_IMPL=chain timeout 300 python bench.py $B > gpurun_out/r2_ab_chain1.json 2> gpurun_out/r2_ab_chain1.err; echo "chain1 rc=$?"
TL_DECODE_IMPL=chain TL_CHAIN_LAYERS=2 timeout 300 python bench.py $B > gpurun_out/r2_ab_chain2.json 2> gpurun_out/r2_ab_chain2.err; echo "chain2 rc=$?"
timeout 1200 python -m pytest tests/test_decode_chain_gpu.py tests/test_parity_scale_gpu.py tests/test_train_gpu.py tests/test_worker_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_gpu_tests2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_gpu_tests2.log
cat gpurun_out/r2_trace_chain.txt | tail -30
for f in kernels chain1 chain2; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_ab_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), 'tok/s', d['gpu_launches'])" 2>&1 | tail -1; done
