#!/usr/bin/env bash
# 1 GPU: tcgen05 attention backward — parity test (bounded), then a timing A/B of the training step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py -q -x -k "attn_bwd" > gpurun_out/r2_gpu_tests19.log 2>&1; echo "pytest attn_bwd rc=$?"; tail -25 gpurun_out/r2_gpu_tests19.log
