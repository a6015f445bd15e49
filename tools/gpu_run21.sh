#!/usr/bin/env bash
mkdir -p gpurun_out
TL_ADAM_STREAM=0 timeout 200 python tools/bench_adam.py 2>&1 | tail -1
TL_ADAM_STREAM=1 timeout 200 python tools/bench_adam.py 2>&1 | tail -1
TL_ADAM_STREAM=0 timeout 200 python tools/bench_adam.py 2>&1 | tail -1
TL_ADAM_STREAM=1 timeout 200 python tools/bench_adam.py 2>&1 | tail -1
