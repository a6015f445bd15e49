"""``PeerRing`` — decode-time hops between pipeline stages through peer-mapped HBM instead of NCCL send/recv.

The reference ships every inter-shard activation as a pickled tensor through its node process
(/root/reference/tensorlink/ml/module.py:438-462, p2p/torch_node.py:195-214, ml/worker.py:297-357).  For single-token
rows that path (and an NCCL send/recv pair with the host in the loop) costs more than the 7 KB it moves, so on one
NVSwitch node every stage owns a *mailbox* in its own HBM, exported with CUDA IPC:

    flags   [2 * n_slots] uint32     sequence numbers written by the neighbour: hidden rows / token ids per slot
    x_in    [n_slots, max_batch, H]  bf16 rows arriving from the previous stage (the stage computes on them in place)
    ids_in  [n_slots, max_batch]     int64 token ids arriving at the first stage from the last one

and maps the mailbox of the next stage (and the last stage that of the first).  The final GEMV of a stage stores
straight into the neighbour's ``x_in`` over NVLink, ``tl_peer_signal`` publishes the hop, and the neighbour's captured
decode graph begins with ``tl_peer_wait`` (csrc/peer.cu).  Send/wait counters are device-resident, so one graph per
(slot, rows) replays for every token and the host enqueues a whole generation without synchronising.
"""
from __future__ import annotations

from typing import List

import torch

from .. import native as nat
from .link import StageLink

_FLAG_BYTES = 256


class PeerRing:
    def __init__(self, link: StageLink, n_slots: int, max_batch: int, hidden: int, max_new: int, device):
        nat.require_device()
        self.link, self.n_slots, self.max_batch, self.hidden, self.device = link, n_slots, max_batch, hidden, torch.device(device)
        assert 2 * n_slots * 4 <= _FLAG_BYTES
        x_bytes = n_slots * max_batch * hidden * 2
        ids_bytes = n_slots * max_batch * 8
        self.nbytes = _FLAG_BYTES + x_bytes + ids_bytes
        self._off_x, self._off_ids = _FLAG_BYTES, _FLAG_BYTES + x_bytes
        self._base, handle = nat.peer_alloc(self.nbytes)
        self._opened: List[int] = []
        mine = self._views(self._base)
        self.flag_x, self.flag_ids, self.x_in, self.ids_in = mine
        handles = link.all_gather_object(handle)
        self.next_flag_x = self.next_x_in = self.first_flag_ids = self.first_ids_in = None
        if not link.last:
            p = nat.peer_open(handles[link.rank + 1]); self._opened.append(p)
            self.next_flag_x, _, self.next_x_in, _ = self._views(p)
        if link.last and link.world > 1:
            p = nat.peer_open(handles[0]); self._opened.append(p)
            _, self.first_flag_ids, _, self.first_ids_in = self._views(p)
        dev = self.device
        # private, device-resident progress counters (one per slot) and the log of generated tokens
        self.want_x = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        self.want_ids = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        self.sent_x = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        self.sent_ids = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.wait_ns = torch.zeros(1, dtype=torch.int64, device=dev)
        self.max_new = max_new
        self.out_log = torch.zeros(n_slots, max_batch, max_new, dtype=torch.int64, device=dev)
        self.step_dev = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        # first use of the mailbox kernels now, sequentially: CUDA's lazy module loading synchronises the context, which
        # must never happen for the first time while a waiter of this process is already spinning
        scratch = torch.zeros(4, dtype=torch.int32, device=dev)
        nat.peer_signal(scratch[0:1], scratch[1:2])
        nat.peer_wait(scratch[0:1], scratch[2:3], self.err, self.wait_ns)
        nat.peer_put(self.x_in[0][0], self.x_in[0][0], scratch[0:1], scratch[1:2])
        torch.cuda.synchronize(dev)
        link.barrier()

    def _views(self, base: int):
        ns, mb, H = self.n_slots, self.max_batch, self.hidden
        raw = nat.tensor_from_ptr(base, self.nbytes)
        flags = raw[:_FLAG_BYTES].view(torch.int32)
        x = raw[self._off_x:self._off_ids].view(torch.bfloat16).view(ns, mb, H)
        ids = raw[self._off_ids:].view(torch.int64).view(ns, mb)
        return flags[:ns], flags[ns:2 * ns], x, ids

    # ---- per-generation reset: every rank clears what it owns, then all ranks meet before the first signal
    def reset(self):
        self.flag_x.zero_(); self.flag_ids.zero_()
        for t in (self.want_x, self.want_ids, self.sent_x, self.sent_ids, self.err, self.wait_ns, self.step_dev):
            t.zero_()
        torch.cuda.synchronize(self.device)
        self.link.barrier()

    # ---- graph-capturable pieces (slot s)
    # ``bump``: a device int32 the same one-thread kernel increments (the stage's KV length on the opening wait, its
    # cache position on the closing signal), so a decode step needs no separate bookkeeping launches
    def wait_x(self, s: int, bump=None):
        nat.peer_wait(self.flag_x[s:s + 1], self.want_x[s:s + 1], self.err, self.wait_ns, bump=bump)

    def wait_ids(self, s: int, bump=None):
        nat.peer_wait(self.flag_ids[s:s + 1], self.want_ids[s:s + 1], self.err, self.wait_ns, bump=bump)

    def signal_x(self, s: int, bump=None):
        nat.peer_signal(self.next_flag_x[s:s + 1], self.sent_x[s:s + 1], bump=bump)

    def signal_ids(self, s: int, bump=None):
        nat.peer_signal(self.first_flag_ids[s:s + 1], self.sent_ids[s:s + 1], bump=bump)

    def log_token(self, s: int, B: int):
        """out_log[s, :B, step] = ids_in[s, :B]; ++step (first stage)."""
        nat.append_token(self.ids_in[s][:B], self.out_log[s][:B], self.step_dev[s:s + 1])

    def check(self):
        if int(self.err.item()):
            raise nat.NativeError("peer mailbox wait timed out (a neighbouring stage stopped signalling)")

    def close(self):
        for p in self._opened:
            nat.peer_close(p)
        self._opened.clear()
        if self._base:
            nat.peer_free(self._base)
            self._base = 0
