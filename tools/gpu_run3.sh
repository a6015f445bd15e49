#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_chain_gpu.py tests/test_sampling_gpu.py tests/test_model_gpu.py tests/test_parity_scale_gpu.py -q -x > gpurun_out/r2_gpu_tests3.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_gpu_tests3.log
timeout 300 python tools/trace_chain.py > gpurun_out/r2_trace_chain_v2.txt 2>&1; echo "trace rc=$?"
tail -28 gpurun_out/r2_trace_chain_v2.txt
TL_CHAIN_DYNAMIC=0 timeout 300 python tools/trace_chain.py > gpurun_out/r2_trace_chain_v2_static.txt 2>&1; echo "trace static rc=$?"
tail -28 gpurun_out/r2_trace_chain_v2_static.txt
B="--no-train --no-cpu-baseline --no-parity-check --steps 2 --warmup 3"
timeout 300 python bench.py $B > gpurun_out/r2_ab_chainv2.json 2> gpurun_out/r2_ab_chainv2.err; echo "chainv2 rc=$?"
TL_CHAIN_DYNAMIC=0 timeout 300 python bench.py $B > gpurun_out/r2_ab_chainv2_static.json 2> gpurun_out/r2_ab_chainv2_static.err; echo "chainv2 static rc=$?"
TL_ADAM_OVERLAP=0 timeout 400 python tools/bench_train.py > gpurun_out/r2_train_noverlap.json 2> gpurun_out/r2_train_noverlap.err; echo "train0 rc=$?"
TL_ADAM_OVERLAP=1 timeout 400 python tools/bench_train.py > gpurun_out/r2_train_overlap.json 2> gpurun_out/r2_train_overlap.err; echo "train1 rc=$?"
for f in ab_chainv2 ab_chainv2_static train_noverlap train_overlap; do python -c "
import json
d=json.loads(open('gpurun_out/r2_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), d['unit'], round(d['ms_per_step'],2), 'ms')" 2>&1 | tail -1; done
