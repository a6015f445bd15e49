#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 120 python -m pytest tests/test_api_gpu.py -q -x -m gpu -k "generate_edge or left_padded" 2>&1 | tail -2
