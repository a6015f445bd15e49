#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -q -x -k "two_gpu" > gpurun_out/r2_gpu_tests23_n2.log 2>&1; echo "pytest n2 rc=$?"; tail -12 gpurun_out/r2_gpu_tests23_n2.log
