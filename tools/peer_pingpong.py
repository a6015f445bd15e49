"""One-way latency of a 7 KB decode hop between two GPUs: peer-mapped mailbox (csrc/peer.cu) vs NCCL send/recv.
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_pingpong.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorlink_b200 import native as nat  # noqa: E402
from tensorlink_b200.p2p.link import StageLink, init_process_group_from_env  # noqa: E402

init_process_group_from_env("nccl")
rank = dist.get_rank()
link = StageLink.from_env()
H, ITERS = 3584, 2000
ptr, handle = nat.peer_alloc(256 + H * 2)
handles = link.all_gather_object(handle)
peer = nat.peer_open(handles[1 - rank])
mine, theirs = nat.tensor_from_ptr(ptr, 256 + H * 2), nat.tensor_from_ptr(peer, 256 + H * 2)
flag, buf = mine[:4].view(torch.int32), mine[256:].view(torch.bfloat16)
pflag, pbuf = theirs[:4].view(torch.int32), theirs[256:].view(torch.bfloat16)
want, sent, err = (torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(3))
src = torch.randn(H, device="cuda").bfloat16()
nat.peer_put(buf, src, flag, sent); nat.peer_wait(flag, want, err)            # load the kernels before anyone spins
torch.cuda.synchronize(); flag.zero_(); want.zero_(); sent.zero_(); torch.cuda.synchronize(); dist.barrier()


def pingpong(n):
    for _ in range(n):
        if rank == 0:
            nat.peer_put(pbuf, src, pflag, sent)
            nat.peer_wait(flag, want, err, None)
        else:
            nat.peer_wait(flag, want, err, None)
            nat.peer_put(pbuf, buf, pflag, sent)


g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    pingpong(100)
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.replay(); torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(ITERS // 100):
    g.replay()
e1.record(); torch.cuda.synchronize()
peer_us = e0.elapsed_time(e1) * 1e3 / (2 * ITERS)
assert int(err) == 0

x = torch.empty(H, dtype=torch.bfloat16, device="cuda")
for it in range(3):
    n = 20 if it < 2 else 500
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(n):
        if rank == 0:
            dist.send(src, 1); dist.recv(x, 1)
        else:
            dist.recv(x, 0); dist.send(x, 0)
    e1.record(); torch.cuda.synchronize()
nccl_us = e0.elapsed_time(e1) * 1e3 / (2 * 500)
if rank == 0:
    print(json.dumps({"payload_bytes": H * 2, "peer_mailbox_one_way_us": peer_us, "nccl_send_recv_one_way_us": nccl_us,
                      "note": "peer: put kernel (copy + release flag) -> acquire-wait kernel, 100 round trips per CUDA graph; "
                              "nccl: blocking dist.send/recv pairs issued from the host"}))
dist.barrier()
dist.destroy_process_group()
