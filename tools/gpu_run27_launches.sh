#!/usr/bin/env bash
# 1 GPU: the first 1200 kernel launches of the bench command on the final tree (same command as the round-1 list)
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --new 4 --no-cpu-baseline --no-train --no-parity-check > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_launches_bench.csv; tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
