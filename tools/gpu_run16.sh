#!/usr/bin/env bash
# 1 GPU: small-model decode (config 2, Qwen2.5-0.5B): GEMV ring small enough for two kernels to co-reside under PDL
mkdir -p gpurun_out
B="--workload cfg2 --no-train --no-parity-check --no-cpu-baseline --steps 2 --warmup 3"
run() { env "$@" timeout 300 python bench.py $B > gpurun_out/r2_cfg2_$TAG.json 2> gpurun_out/r2_cfg2_$TAG.err; echo "$TAG rc=$?"; }
TAG=base run TL_X=0
TAG=ring100 run TL_GEMV_RING_KB=100
TAG=ring72 run TL_GEMV_RING_KB=72
TAG=ring100_pdlattn run TL_GEMV_RING_KB=100 TL_PDL_ATTN=1
TAG=ring100_pf0 run TL_GEMV_RING_KB=100 TL_PREFETCH_MB=0
TAG=k2 run TL_GEMV_CTAS_PER_SM=2
for f in base ring100 ring72 ring100_pdlattn ring100_pf0 k2; do python -c "
import json
d=json.loads(open('gpurun_out/r2_cfg2_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], d['roofline']['decode_step'].get('decode_only'))" 2>&1 | tail -1; done
