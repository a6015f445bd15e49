"""One pipeline stage on one B200: parameters + per-micro-batch shard operators + captured decode graphs.

This is the worker half of the reference's hot path (/root/reference/tensorlink/ml/worker.py):
``load_module`` (:452-505) -> ``CudaStage.__init__``; ``_handle_forward`` (:297-357) -> ``prefill`` / ``decode``;
``_handle_generate`` (:359-441) -> the decode graph.  There is no polling loop and no IPC: the stage is driven
in-process by ``DistributedModel`` and neighbours are reached through ``StageLink``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import native as nat
from .configs import ShardModelConfig
from .shard import CudaLayerGroup, ShardParams, gemv_max_rows


class CudaStage:
    def __init__(self, cfg: ShardModelConfig, layer_ids, has_embed: bool, has_head: bool, device,
                 max_batch: int, max_seq: int, n_slots: int = 1, training: bool = False,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 1234, init: str = "seeded",
                 max_tokens: Optional[int] = None):
        nat.require_device()
        self.cfg = cfg
        self.device = torch.device(device)
        self.has_embed, self.has_head = has_embed, has_head
        self.supports_training, self.trainer = bool(training), None
        self.params = ShardParams(cfg, layer_ids, has_embed, has_head, self.device, with_grad=training)
        if state_dict is not None:
            self.params.load_hf_state_dict(state_dict)
        elif init == "device":
            self.params.init_on_device(seed)
        else:
            self.params.init_seeded(seed)
        self.max_batch, self.max_seq = max_batch, max_seq
        self.slots: List[CudaLayerGroup] = [CudaLayerGroup(cfg, self.params, max_batch, max_seq, max_tokens)
                                            for _ in range(n_slots)]
        dev = self.device
        # decode-time fixed buffers (graph inputs/outputs), one set per slot
        self.x_dec = [torch.zeros(max_batch, cfg.hidden, dtype=torch.bfloat16, device=dev) for _ in range(n_slots)]
        self.ids_dec = [torch.zeros(max_batch, dtype=torch.int64, device=dev) for _ in range(n_slots)]
        self.graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        if has_head:
            self.head_ws = torch.empty(max(nat.lmhead_ws(min(max_batch, 8), cfg.vocab), max_batch * 64 * 8 + 256),
                                       dtype=torch.uint8, device=dev)
            self.logits_dec = torch.empty(max_batch, cfg.vocab, dtype=torch.bfloat16, device=dev)
            self.hn = torch.empty(max_batch, cfg.hidden, dtype=torch.bfloat16, device=dev)
            self.sampling: Optional[dict] = None        # set by generate(do_sample=True): temperature / top_k / top_p / seed
            self.sample_ctr = torch.zeros(n_slots, max_batch, dtype=torch.int32, device=dev)
            self.sample_ws: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------------ pieces
    def embed(self, ids: torch.Tensor) -> torch.Tensor:
        """[B,S] int64 -> [B,S,H]  (host-side ``embed_tokens`` in the reference, module.py:1023-1056)."""
        return nat.embed_fwd(ids.contiguous(), self.params.v["embed"])

    def prefill(self, hidden: torch.Tensor, past_len: int = 0, slot: int = 0) -> torch.Tensor:
        return self.slots[slot].prefill(hidden, past_len)

    def head_logits(self, hidden: torch.Tensor) -> torch.Tensor:
        """final norm + lm_head over [N,H] -> bf16 logits [N,V]."""
        cfg, v = self.cfg, self.params.v
        n = hidden.shape[0]
        if n <= gemv_max_rows():
            return nat.gemv(hidden.contiguous(), v["head"], norm_w=v["norm"], eps=cfg.rms_eps)
        hn = nat.rmsnorm_fwd(hidden.contiguous(), v["norm"], cfg.rms_eps)
        return nat.gemm(hn, v["head"])

    def set_sampling(self, sampling: Optional[dict]):
        """None = greedy.  Changing the mode drops the captured decode graphs (the launch sequence differs)."""
        if not self.has_head or sampling == self.sampling:
            return
        self.sampling = sampling
        self.graphs.clear()
        self.sample_ctr.zero_()
        if sampling is not None and self.sample_ws is None:
            self.sample_ws = torch.empty(nat.sample_ws(self.max_batch), dtype=torch.uint8, device=self.device)

    def head_argmax(self, hidden: torch.Tensor, ids_out: torch.Tensor, slot: int = 0):
        """next token for [B,H] rows -> ids_out [B] int64: greedy (bit-exact target: torch.argmax of bf16 logits) or, after
        ``set_sampling``, one draw per row from the warped distribution (csrc/sample.cu)."""
        self._head_greedy(hidden, ids_out)
        if self.sampling is not None:
            B = hidden.shape[0]
            s = self.sampling
            nat.sample(self.logits_dec[:B], ids_out, self.sample_ctr[slot], self.sample_ws, s["temperature"], s["top_k"], s["top_p"],
                       s["seed"] + 0x9E3779B97F4A7C15 * slot)

    def _head_greedy(self, hidden: torch.Tensor, ids_out: torch.Tensor):
        cfg, v = self.cfg, self.params.v
        B = hidden.shape[0]
        if B <= gemv_max_rows():
            nat.lmhead_argmax(hidden, v["head"], v["norm"], cfg.rms_eps, ids_out, self.logits_dec[:B], self.head_ws)
        else:
            nat.rmsnorm_fwd(hidden, v["norm"], cfg.rms_eps, out=self.hn[:B])
            nat.gemm(self.hn[:B], v["head"], out=self.logits_dec[:B])
            nat.argmax_bf16(self.logits_dec[:B], ids_out, self.head_ws)

    # ------------------------------------------------------------------------------------------ decode step
    def _decode_body_ring(self, slot: int, B: int, ring):
        """The decode step of a multi-stage pipeline with the hops on peer memory (p2p/peer.py): wait for this slot's
        input in the local mailbox, run the layers, let the last kernel store into the neighbour's mailbox, signal."""
        grp = self.slots[slot]
        if self.has_embed:
            ring.wait_ids(slot, bump=grp.kvlen_dev)              # the opening wait also counts the new key in
            ring.log_token(slot, B)
            x = self.x_dec[slot][:B]
            nat.embed_fwd(ring.ids_in[slot][:B], self.params.v["embed"], out=x)
        else:
            ring.wait_x(slot, bump=grp.kvlen_dev)
            x = ring.x_in[slot][:B]
        if self.has_head:
            grp.decode_step_inplace(x, advance=False)
            self.head_argmax(x, ring.first_ids_in[slot][:B], slot)
            ring.signal_ids(slot, bump=grp.pos_dev)              # the closing signal also advances the cache position
        else:
            grp.decode_step_inplace(x, out=ring.next_x_in[slot][:B], advance=False)
            ring.signal_x(slot, bump=grp.pos_dev)

    def _decode_body(self, slot: int, B: int, ring=None):
        if ring is not None:
            self._decode_body_ring(slot, B, ring)
            return
        x = self.x_dec[slot][:B]
        if self.has_embed:
            nat.embed_fwd(self.ids_dec[slot][:B], self.params.v["embed"], out=x)
        self.slots[slot].decode_step_inplace(x)
        if self.has_head:
            self.head_argmax(x, self.ids_dec[slot][:B], slot)

    def decode(self, slot: int, B: int, use_graph: bool = True, ring=None):
        """One token for slot's rows: [embed ->] layers [-> norm + lm_head + argmax], as ONE graph launch.
        Inputs/outputs are the fixed buffers ``ids_dec[slot]`` / ``x_dec[slot]``, or the mailboxes of ``ring``."""
        if not use_graph:
            self._decode_body(slot, B, ring)
            return
        key = (slot, B) if ring is None else (slot, B, id(ring))
        g = self.graphs.get(key)
        if g is None:
            # warm up outside capture (first-use attribute setting, tensor-map cache), restoring the state it touches
            grp = self.slots[slot]
            saved = (grp.pos_dev.clone(), grp.kvlen_dev.clone(), self.ids_dec[slot].clone(), self.x_dec[slot].clone())
            ctr_saved = self.sample_ctr.clone() if self.has_head else None      # the warm-up step must not consume a draw
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._decode_body(slot, B)
            torch.cuda.current_stream().wait_stream(side)
            grp.pos_dev.copy_(saved[0]); grp.kvlen_dev.copy_(saved[1])
            self.ids_dec[slot].copy_(saved[2]); self.x_dec[slot].copy_(saved[3])
            if ctr_saved is not None:
                self.sample_ctr.copy_(ctr_saved)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._decode_body(slot, B, ring)      # (the warm-up above never touches the mailboxes)
            # capture does not execute; state is still the saved one
            self.graphs[key] = g
        g.replay()

    def check(self):
        for g in self.slots:
            g.check()

    def n_decode_launches(self, B: int, ring: bool = False) -> int:
        """Kernel launches inside one decode step of this stage (for bench.py's gpu_launches claim)."""
        fused = self.slots[0].T_max <= self.slots[0].FUSED_DECODE_MAX_T
        gemv = B <= gemv_max_rows()
        if self.slots[0].chain_ok(B):
            n = self.slots[0].n_chain_launches() + 2       # qkv of the first layer + one persistent launch per layer group
        elif self.slots[0].dq_ok(B):
            # first qkv + per layer: attention (1 fused / 3), o, gate/up, [down + next qkv] as one chain launch
            n = 1 + len(self.slots[0].layer_ids) * (3 + (1 if fused else 3)) + 2
        else:
            # GEMV path: 4 Linears + attention (1 fused / 3); batched: 4 GEMMs + 3 split-K reduce(+norm) passes + attention
            n = len(self.slots[0].layer_ids) * ((7 if gemv else 10) - (2 if fused else 0)) + 2 + (0 if gemv else 1)
        if ring:
            n += (3 if self.has_embed else 2) - 2    # wait (+ token log) + signal, which also do the two position updates
        if self.has_embed:
            n += 1
        if self.has_head:
            n += 3 if B <= gemv_max_rows() else 4
        return n
