#!/usr/bin/env bash
# 1 GPU: auto 2-CTA/SM GEMV for small weights — config 2 A/B, then the round-end sequence (full GPU suite, bench, smoke)
mkdir -p gpurun_out
B="--workload cfg2 --no-train --no-parity-check --no-cpu-baseline --steps 2 --warmup 3"
run() { env "$@" timeout 300 python bench.py $B > gpurun_out/r2_cfg2_$TAG.json 2> gpurun_out/r2_cfg2_$TAG.err; echo "$TAG rc=$?"; }
TAG=auto run TL_X=0
TAG=auto_pdlattn run TL_PDL_ATTN=1
TAG=k1 run TL_GEMV_CTAS_PER_SM=1
for f in auto auto_pdlattn k1; do python -c "
import json
d=json.loads(open('gpurun_out/r2_cfg2_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], d['roofline']['decode_step'].get('decode_only'))" 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2_gpu_tests17.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2_gpu_tests17.log
timeout 900 python bench.py > gpurun_out/r2_final2_n1.json 2> gpurun_out/r2_final2_n1.err; echo "bench rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final2_smoke.log 2>&1; echo "smoke rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final2_n1.json').read().strip().splitlines()[-1])
print('final', d.get('value'), d.get('unit'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'train', (d.get('train') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
