"""Prefill attention: legacy mma.sync tiles vs the tcgen05 kernel, CUDA-event timed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorlink_b200 import native as nat  # noqa: E402

for (B, S, n_h, n_kv, d) in ((8, 512, 28, 4, 128), (1, 4096, 28, 4, 128), (8, 512, 14, 2, 64), (16, 1024, 32, 8, 128)):
    q = torch.randn(B, S, n_h, d, device="cuda").bfloat16()
    kc = torch.randn(B, n_kv, S, d, device="cuda").bfloat16()
    vc = torch.randn(B, n_kv, S, d, device="cuda").bfloat16()
    out = torch.empty(B, S, n_h * d, dtype=torch.bfloat16, device="cuda")
    flops = 4.0 * B * S * S * n_h * d / 2
    res = {"B": B, "S": S, "n_h": n_h, "n_kv": n_kv, "d": d}
    outs = {}
    for impl in ("mma", "tc"):
        os.environ["TL_ATTN_IMPL"] = impl
        for _ in range(3):
            nat.attn_prefill_fwd(q, kc, vc, out, None, B, S, 0, n_h, n_kv, d, d ** -0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            nat.attn_prefill_fwd(q, kc, vc, out, None, B, S, 0, n_h, n_kv, d, d ** -0.5)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        res[impl + "_us"] = us
        res[impl + "_tflops"] = flops / us / 1e6
        outs[impl] = out.float().clone()
    res["rel_diff"] = float((outs["tc"] - outs["mma"]).norm() / outs["mma"].norm())
    print(json.dumps(res))
