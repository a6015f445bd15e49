#!/usr/bin/env python
"""Timeline of the decode-chain kernel inside the captured decode graph (device-side globaltimer stamps).

    python tools/trace_chain.py [--model Qwen/Qwen2.5-7B] [--prompt 32] [--steps 8]

For every chain launch of one decode step (CTA 0 and the last CTA): when each job started, had its input staged, finished
its work and passed its dependency — averaged over the layers — plus the gap between consecutive launches.  Numbers in
microseconds.  This is a measurement aid; nothing here is a bench value."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorlink_b200 import native as nat  # noqa: E402
from tensorlink_b200.ml import DistributedModel  # noqa: E402
from tensorlink_b200.ml.configs import get_config  # noqa: E402
from tensorlink_b200.ml.weights import synthetic_tokens  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Qwen/Qwen2.5-7B")
    ap.add_argument("--prompt", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--rows", type=int, default=1)
    args = ap.parse_args()
    cfg = get_config(args.model)
    torch.cuda.set_device(0)
    dm = DistributedModel(args.model, training=False, max_batch=args.rows, max_seq=args.prompt + args.steps + 16, init="device",
                          max_tokens=args.rows * args.prompt)
    st = dm.stage
    grp = st.slots[0]
    n_launch = grp.n_chain_launches() - 1
    ids = synthetic_tokens(cfg, args.rows, args.prompt).cuda()
    x = st.prefill(st.embed(ids), 0, 0)
    st.head_argmax(x[:, -1, :].contiguous(), st.ids_dec[0][:args.rows])
    trace = torch.zeros(4 * n_launch, nat.CHAIN_TRACE_WORDS, dtype=torch.int64, device="cuda")
    nat.decode_chain_trace(trace)
    for _ in range(args.steps):                       # first call: eager warm-up (slots 0..n-1) + capture (slots n..2n-1)
        st.decode(0, args.rows)
    torch.cuda.synchronize()
    nat.decode_chain_trace(None)
    raw = trace[n_launch:2 * n_launch].cpu()
    head = 2 * (nat.CHAIN_MAX_JOBS + 1) * 4
    t = raw[:, :head].reshape(n_launch, 2, nat.CHAIN_MAX_JOBS + 1, 4).double() * 1e-3   # us
    done_all = raw[:, head:head + nat.CHAIN_MAX_JOBS * 160].reshape(n_launch, nat.CHAIN_MAX_JOBS, 160).double() * 1e-3
    stat = raw[:, head + nat.CHAIN_MAX_JOBS * 160:].reshape(n_launch, nat.CHAIN_MAX_JOBS, 4).double()
    names = ["attn", "o", "gate_up", "down", "qkv_next"]
    print(f"{args.model}: {n_launch} chain launches per step, last replay; times in us relative to each launch's kernel entry (CTA 0)")
    for sel, who in ((0, "CTA 0"), (1, "last CTA")):
        print(f"--- {who}")
        base = t[:, 0, nat.CHAIN_MAX_JOBS, 0]
        print("  kernel entry -> previous grid done (griddepcontrol.wait): %.2f   kernel entry -> exit: %.2f" % (
            float((t[1:-1, sel, -1, 1] - base[1:-1]).mean()), float((t[1:-1, sel, -1, 2] - base[1:-1]).mean())))
        for j, nm in enumerate(names):
            rows = t[1:-1, sel, j]                     # skip the first / last launch (different neighbours)
            if float(rows[:, 0].min()) == 0:
                continue
            rows = rows.clone()
            rows[:, 1] = torch.where(rows[:, 1] == 0, rows[:, 0], rows[:, 1])      # the attention job has no staging stamp
            rel = rows - base[1:-1, None]
            m = rel.mean(0)
            print(f"  {nm:9s} start {m[0]:7.2f}  staged {m[1]:7.2f} (+{m[1] - m[0]:5.2f})  done {m[2]:7.2f} (+{m[2] - m[1]:6.2f})  "
                  f"dep passed {m[3]:7.2f} (+{m[3] - m[2]:5.2f})")
    print("--- all CTAs: spread of the 'work done' time per job (us): last - first, last - median")
    for j, nm in enumerate(names):
        d = done_all[1:-1, j]
        d = d[:, (d > 0).all(0)]
        if d.numel() == 0:
            continue
        print(f"  {nm:9s} {float((d.max(1).values - d.min(1).values).mean()):6.2f} {float((d.max(1).values - d.median(1).values).mean()):6.2f}"
              f"   dep passed - last done: {float((t[1:-1, 0, j, 3] - d.max(1).values).mean()):6.2f}")
    print("--- CTA 0, cycles per job: consumer warp 0 waiting for weights / in the job loop; producer waiting for free ring slots")
    for j, nm in enumerate(names):
        c = stat[1:-1, j].mean(0)
        if float(c[1]) > 0:
            print(f"  {nm:9s} consumer wait {c[0]:9.0f} of {c[1]:9.0f} ({100 * c[0] / c[1]:5.1f} %)   producer wait {c[2]:9.0f}")
    ent = t[:, 0, -1, 0]
    ext = t[:, :, -1, 2].max(1).values
    print("launch period (entry to next entry): %.2f us; exit of launch i -> entry of launch i+1: %.2f us" % (
        float((ent[1:] - ent[:-1]).mean()), float((ent[1:] - ext[:-1]).mean())))


if __name__ == "__main__":
    main()
