#!/usr/bin/env bash
mkdir -p gpurun_out
for cfg in "16 192" "16 0" "8 192" "8 0"; do
  set -- $cfg
  TL_CHAIN_STAGE_KB=$1 TL_CHAIN_L2_AHEAD_KB=$2 timeout 300 python tools/trace_chain.py > gpurun_out/r2_trace_v3_s$1_a$2.txt 2>&1; echo "trace $1 $2 rc=$?"
  tail -31 gpurun_out/r2_trace_v3_s$1_a$2.txt
done
B="--no-train --no-cpu-baseline --no-parity-check --steps 2 --warmup 3"
for cfg in "16 192" "16 0" "8 192" "8 512"; do
  set -- $cfg
  TL_CHAIN_STAGE_KB=$1 TL_CHAIN_L2_AHEAD_KB=$2 timeout 300 python bench.py $B > gpurun_out/r2_ab_v3_s$1_a$2.json 2> gpurun_out/r2_ab_v3_s$1_a$2.err; echo "bench $1 $2 rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/r2_ab_v3_s$1_a$2.json').read().strip().splitlines()[-1]); print('stage $1 KB ahead $2 KB:', round(d['value'],2), d['unit'])" 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_decode_chain_gpu.py tests/test_sampling_gpu.py tests/test_api_gpu.py -q -x > gpurun_out/r2_gpu_tests4.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_gpu_tests4.log
