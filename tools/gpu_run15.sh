#!/usr/bin/env bash
# 1 GPU: fused decode-attention reduce A/B at configs 5 / 3, then the round-end sequence (full GPU suite, bench both arms, smoke)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn_decode" > gpurun_out/r2_gpu_tests15a.log 2>&1; echo "pytest attn rc=$?"; tail -3 gpurun_out/r2_gpu_tests15a.log
B="--no-train --no-parity-check --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 python bench.py --workload cfg5 $B > gpurun_out/r2_bench_cfg5_fused.json 2> gpurun_out/r2_bench_cfg5_fused.err; echo "cfg5 fused rc=$?"
TL_DECODE_ATTN=mma2 timeout 600 python bench.py --workload cfg5 $B > gpurun_out/r2_bench_cfg5_two.json 2> gpurun_out/r2_bench_cfg5_two.err; echo "cfg5 two rc=$?"
timeout 600 python bench.py --workload cfg3 $B > gpurun_out/r2_bench_cfg3_fused.json 2> gpurun_out/r2_bench_cfg3_fused.err; echo "cfg3 fused rc=$?"
TL_DECODE_ATTN=mma2 timeout 600 python bench.py --workload cfg3 $B > gpurun_out/r2_bench_cfg3_two.json 2> gpurun_out/r2_bench_cfg3_two.err; echo "cfg3 two rc=$?"
for f in cfg5_fused cfg5_two cfg3_fused cfg3_two; do python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], d['roofline']['decode_step'].get('decode_only'))" 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2_gpu_tests15.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2_gpu_tests15.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref arm rc=$?"
timeout 900 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; echo "bench rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"
python - <<'PY'
import json
for f in ('r2_final_ref','r2_final_n1'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('unit'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'train', (d.get('train') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, e)
PY
