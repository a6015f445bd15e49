#!/usr/bin/env python
"""The training half of bench.py alone (same code path: ``bench.measure_training``), for A/B runs.

    [torchrun ...] python tools/bench_train.py [--gpus N] [--steps K] [--train-mb-per-stage M] [--train-model ...]
Prints the ``train`` block bench.py would put on its line (one JSON line on rank 0)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--train-model", default="Qwen/Qwen2.5-7B")
    ap.add_argument("--train-batch", type=int, default=8)
    ap.add_argument("--train-seq", type=int, default=512)
    ap.add_argument("--train-mb-per-stage", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true", default=True)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tensorlink_b200.p2p.link import init_process_group_from_env
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        init_process_group_from_env("nccl")
    else:
        torch.cuda.set_device(0)
    _, tf_peak, kind = bench.measured_peaks()
    res = bench.measure_training(args, max(args.gpus, 1), rank, world, tf_peak, kind)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
