"""One training step (fwd + bwd + Adam) of the bench's training workload for ncu / launch-list analysis."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorlink_b200.ml import DistributedModel  # noqa: E402
from tensorlink_b200.ml.configs import get_config  # noqa: E402
from tensorlink_b200.ml.weights import synthetic_tokens  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="Qwen/Qwen2.5-0.5B")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seq", type=int, default=512)
a = ap.parse_args()
cfg = get_config(a.model)
dm = DistributedModel(a.model, training=True, max_batch=a.batch, max_seq=a.seq, init="device", optimizer=torch.optim.Adam, max_tokens=8)
opt = dm.create_optimizer(lr=1e-4)
ids = synthetic_tokens(cfg, a.batch, a.seq).cuda()
for i in range(2):
    if i == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    opt.zero_grad()
    out = dm(ids, labels=ids)
    out.loss.backward()
    opt.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(float(out.loss.detach()))
