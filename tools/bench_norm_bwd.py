"""rmsnorm_bwd timing at training shapes (CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorlink_b200 import native as nat  # noqa: E402

for rows, H in ((4096, 3584), (4096, 896), (8192, 4096)):
    x = torch.randn(rows, H, device="cuda").bfloat16()
    dy = torch.randn(rows, H, device="cuda").bfloat16()
    add = torch.randn(rows, H, device="cuda").bfloat16()
    w = torch.randn(H, device="cuda").bfloat16()
    rstd = torch.rand(rows, device="cuda") + 0.5
    dx = torch.empty_like(x)
    acc = torch.zeros(H, dtype=torch.float32, device="cuda")
    for _ in range(3):
        nat.rmsnorm_bwd(x, w, dy, rstd, dx, acc, dx_add=add)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nat.rmsnorm_bwd(x, w, dy, rstd, dx, acc, dx_add=add)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    print(json.dumps({"rows": rows, "H": H, "us": us, "GBps": rows * H * 2 * 4 / us / 1e3}))
