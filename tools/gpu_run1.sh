#!/usr/bin/env bash
# GPU run 1 of round 2: full GPU test suite + the bench line + cfg3 / cfg5 workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -k "not multigpu" > gpurun_out/r2_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-train --no-parity-check > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err
echo "cfg3 rc=$?"
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-train --no-parity-check --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err
echo "cfg5 rc=$?"
tail -5 gpurun_out/r2_gpu_tests.log
