#!/usr/bin/env bash
mkdir -p gpurun_out
B="--no-train --no-cpu-baseline --no-parity-check --steps 2 --warmup 3"
timeout 300 python bench.py $B > gpurun_out/r2_ab_wide.json 2> gpurun_out/r2_ab_wide.err; echo "wide rc=$?"
TL_GEMV_WIDE=0 timeout 300 python bench.py $B > gpurun_out/r2_ab_narrow.json 2> gpurun_out/r2_ab_narrow.err; echo "narrow rc=$?"
TL_PDL_ATTN=1 timeout 300 python bench.py $B > gpurun_out/r2_ab_wide_pdlattn.json 2> gpurun_out/r2_ab_wide_pdlattn.err; echo "pdlattn rc=$?"
for f in ab_wide ab_narrow ab_wide_pdlattn; do python -c "
import json
d=json.loads(open('gpurun_out/r2_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), d['unit'], d['roofline']['all_gemv_launches']['per_shape_GBps'])" 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -q -k "not multigpu" > gpurun_out/r2_gpu_tests6.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_gpu_tests6.log
