#!/usr/bin/env python
"""Attention backward alone: tcgen05 kernels vs the mma.sync kernels (TL_ATTN_BWD=mma), CUDA events, TFLOP/s.

    python tools/bench_attn_bwd.py            # Qwen2.5-7B head geometry (28 / 4 heads of 128) at training shapes
Causal FLOPs counted: forward 4*B*H*S^2*d/2, backward 2.5x that (5 products of the forward's 2)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorlink_b200 import native as nat  # noqa: E402


def run(B, S, n_h, n_kv, d, impl, iters=20):
    os.environ["TL_ATTN_BWD"] = impl
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(B, S, n_h, d, device="cuda", generator=g) * 0.7).bfloat16()
    kc = (torch.randn(B, n_kv, S, d, device="cuda", generator=g) * 0.7).bfloat16()
    vc = torch.randn(B, n_kv, S, d, device="cuda", generator=g).bfloat16()
    do = torch.randn(B, S, n_h * d, device="cuda", generator=g).bfloat16()
    out = torch.empty(B, S, n_h * d, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(B, n_h, S, dtype=torch.float32, device="cuda")
    nat.attn_prefill_fwd(q, kc, vc, out, lse, B, S, 0, n_h, n_kv, d, d ** -0.5)
    dq = torch.empty_like(q)
    dk = torch.empty(B, n_h, S, d, dtype=torch.bfloat16, device="cuda")
    dv = torch.empty_like(dk)
    ws = torch.empty(nat.attn_bwd_ws(B, S, n_h), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        nat.attn_bwd(q, kc, vc, out, do, lse, dq, dk, dv, ws, B, S, n_h, n_kv, d, d ** -0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        nat.attn_bwd(q, kc, vc, out, do, lse, dq, dk, dv, ws, B, S, n_h, n_kv, d, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.5 * 4.0 * B * n_h * S * S * d / 2
    return {"B": B, "S": S, "heads": f"{n_h}/{n_kv}x{d}", "impl": impl, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
            "dq_sum": float(dq.float().abs().sum())}


if __name__ == "__main__":
    for shape in ((8, 512, 28, 4, 128), (4, 1024, 32, 8, 128), (1, 4096, 28, 4, 128), (8, 512, 14, 2, 64)):
        for impl in ("mma", "tc"):
            print(json.dumps(run(*shape, impl)))
