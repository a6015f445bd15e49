#!/usr/bin/env bash
# 1 GPU: launch list (per-launch durations) of one Qwen2.5-7B training step on the final tree
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_train.csv python tools/profile_train.py --model Qwen/Qwen2.5-7B > gpurun_out/r2_launches_train.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_launches_train.csv; tail -2 gpurun_out/r2_launches_train.log
