#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn_decode" > gpurun_out/r2_gpu_tests10.log 2>&1; echo "pytest attn rc=$?"; tail -4 gpurun_out/r2_gpu_tests10.log
B="--no-train --no-parity-check --steps 2 --warmup 3"
timeout 600 python bench.py --workload cfg5 $B --no-cpu-baseline > gpurun_out/r2_bench_cfg5_mma.json 2> gpurun_out/r2_bench_cfg5_mma.err; echo "cfg5 mma rc=$?"
TL_DECODE_ATTN=simt timeout 600 python bench.py --workload cfg5 $B --no-cpu-baseline > gpurun_out/r2_bench_cfg5_simt.json 2> gpurun_out/r2_bench_cfg5_simt.err; echo "cfg5 simt rc=$?"
timeout 600 python bench.py --workload cfg3 $B > gpurun_out/r2_bench_cfg3_k.json 2> gpurun_out/r2_bench_cfg3_k.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --workload cfg2 $B --no-cpu-baseline > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
for f in cfg5_mma cfg5_simt cfg3_k cfg2; do python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], d['roofline']['decode_step'].get('decode_only'))" 2>&1 | tail -1; done
P="python tools/profile_decode.py --model Qwen/Qwen2.5-7B-Instruct --rows 32 --prompt 1024 --max-seq 4096 --new 3"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_decode_mma" -s 28 -c 2 -o gpurun_out/r2_prof_attn_decode_mma $P > gpurun_out/r2_ncu_attn_mma.log 2>&1; echo "ncu mma rc=$?"
TL_DECODE_ATTN=simt timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_decode_split" -s 28 -c 2 -o gpurun_out/r2_prof_attn_decode_simt $P > gpurun_out/r2_ncu_attn_simt.log 2>&1; echo "ncu simt rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -k "not multigpu" > gpurun_out/r2_gpu_tests10_full.log 2>&1; echo "pytest full rc=$?"
tail -6 gpurun_out/r2_gpu_tests10_full.log
