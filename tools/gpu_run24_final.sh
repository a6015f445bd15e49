#!/usr/bin/env bash
# 1 GPU: the round-end sequence on the final tree — full GPU suite, bench (reference arm, then own arm), smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2_gpu_tests_final.log
timeout 600 python bench.py --impl reference > gpurun_out/r2_final3_ref.json 2> gpurun_out/r2_final3_ref.err; echo "ref arm rc=$?"
timeout 600 python bench.py > gpurun_out/r2_final3_n1.json 2> gpurun_out/r2_final3_n1.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final3_smoke.log 2>&1; echo "smoke rc=$?"
python - <<'PY'
import json
for f in ('r2_final3_ref','r2_final3_n1'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('unit'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'train', (d.get('train') or {}).get('value'), (d.get('train') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, e)
PY
