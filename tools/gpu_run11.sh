#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -q -x -k "attn_decode or upstream or other_optimizer" > gpurun_out/r2_gpu_tests11.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests11.log
B="--no-train --no-parity-check --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 python bench.py --workload cfg5 $B > gpurun_out/r2_bench_cfg5_mma2.json 2> gpurun_out/r2_bench_cfg5_mma2.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg3 $B > gpurun_out/r2_bench_cfg3_mma2.json 2> gpurun_out/r2_bench_cfg3_mma2.err; echo "cfg3 rc=$?"
TL_DECODE_IMPL=chain TL_CHAIN_LAYERS=3 timeout 600 python bench.py --workload cfg2 $B > gpurun_out/r2_bench_cfg2_chain3.json 2> gpurun_out/r2_bench_cfg2_chain3.err; echo "cfg2 chain3 rc=$?"
TL_DECODE_IMPL=chain timeout 600 python bench.py --workload cfg2 $B > gpurun_out/r2_bench_cfg2_chain1.json 2> gpurun_out/r2_bench_cfg2_chain1.err; echo "cfg2 chain1 rc=$?"
for f in cfg5_mma2 cfg3_mma2 cfg2_chain3 cfg2_chain1; do python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], d['roofline']['decode_step'].get('decode_only'))" 2>&1 | tail -1; done
P="python tools/profile_decode.py --model Qwen/Qwen2.5-7B-Instruct --rows 32 --prompt 1024 --max-seq 4096 --new 3"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_decode_mma" -s 28 -c 2 -o gpurun_out/r2_prof_attn_decode_mma2 $P > gpurun_out/r2_ncu_attn_mma2.log 2>&1; echo "ncu mma rc=$?"
