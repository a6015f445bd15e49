"""A few GEMM launches for ncu (tensor-pipe utilisation, stall reasons)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorlink_b200 import native as nat  # noqa: E402

shapes = [(4096, 9728, 896, 0), (4096, 4608, 3584, 0), (8192, 8192, 8192, 0)]
nat.require_device()
for (M, N, K, fl) in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    o = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        nat.gemm(a, w, out=o, flags=fl)
torch.cuda.synchronize()
