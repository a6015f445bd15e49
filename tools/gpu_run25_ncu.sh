#!/usr/bin/env bash
# 1 GPU: ncu --set full of the two tcgen05 attention-backward kernels (first launch of each: B=8, S=512, 28/4 heads of 128)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_d.*_tc_kernel" -c 2 -f -o gpurun_out/r2_prof_attn_bwd_tc python tools/bench_attn_bwd.py > gpurun_out/r2_ncu_attn_bwd_tc.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_attn_bwd_tc.log; ls -la gpurun_out/r2_prof_attn_bwd_tc.ncu-rep
