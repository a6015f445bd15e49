#!/usr/bin/env bash
# ncu evidence for the batched-decode path of BASELINE config 5 (32 rows, KV cache sized for 4096): launch list + --set full
mkdir -p gpurun_out
P="python tools/profile_decode.py --model Qwen/Qwen2.5-7B-Instruct --rows 32 --prompt 1024 --max-seq 4096 --new 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_cfg5.csv $P > gpurun_out/r2_ncu_cfg5_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_decode_kernel|attn_decode_reduce" -s 8 -c 4 -o gpurun_out/r2_prof_attn_decode $P > gpurun_out/r2_ncu_attn.log 2>&1; echo "attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16_kernel|splitk_reduce" -s 0 -c 8 -o gpurun_out/r2_prof_gemm_m32 $P > gpurun_out/r2_ncu_gemm.log 2>&1; echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
