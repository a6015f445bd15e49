#!/usr/bin/env bash
# 2 GPUs: multi-rank tests (EOS early stop on both transports), API / training / model tests after the generate-loop changes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/r2_gpu_tests18_n2.log 2>&1; echo "pytest n2 rc=$?"; tail -5 gpurun_out/r2_gpu_tests18_n2.log
CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests/test_api_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py tests/test_sampling_gpu.py tests/test_worker_gpu.py -q -x -m gpu > gpurun_out/r2_gpu_tests18.log 2>&1; echo "pytest 1gpu rc=$?"; tail -4 gpurun_out/r2_gpu_tests18.log
