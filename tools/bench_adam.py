#!/usr/bin/env python
"""AdamW sweep alone: GB/s over a 2e9-parameter arena (22 bytes of HBM traffic per parameter), CUDA events."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorlink_b200 import native as nat  # noqa: E402

if __name__ == "__main__":
    n = 2_000_000_000
    p = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    g = torch.full((n,), 0.01, dtype=torch.bfloat16, device="cuda")
    m = torch.zeros(n, dtype=torch.float32, device="cuda")
    v = torch.zeros(n, dtype=torch.float32, device="cuda")
    for t in range(1, 3):
        nat.adamw_step(p, g, m, v, 1e-4, 0.9, 0.999, 1e-8, 0.0, t, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(3, 8):
        nat.adamw_step(p, g, m, v, 1e-4, 0.9, 0.999, 1e-8, 0.0, t, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"stream_hint": os.environ.get("TL_ADAM_STREAM", "0"), "ms": round(ms, 3), "GBps": round(22.0 * n / ms / 1e6, 1)}))
