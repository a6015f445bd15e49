"""The CUDA shard operator: a contiguous layer range executed by the sm_100a kernels.

Mirrors the reference's shard operator ``LayerGroupModule`` (/root/reference/tensorlink/ml/injector.py:154-281):
same constructor role (a list of layers + the loop live-ins), ``num_layers`` attribute, ``forward(**kwargs) ->
dict`` returning ``kwargs ∪ {hidden_states}``.  What differs is everything underneath: instead of ``exec``-ing
HF's loop body over ``nn.Module`` layers, each layer is five to eight kernel launches on a resident bf16
parameter arena with a resident KV cache; masks, RoPE tables and positions are generated on the device and
never cross a shard boundary (the reference ships them with every call, injector.py:508-556).

HBM layout (per shard):
  params   one flat bf16 arena; per layer  ln1 | wqkv[(n_h+2n_kv)d, H] | bqkv | (q_norm,k_norm) | wo[H, n_h d] |
           ln2 | wgu[2I, H] (row 2j = gate_j, row 2j+1 = up_j) | wd[H, I]; then embed / final norm / lm_head
           on the ranks that own them.  Every tensor starts on a 256-byte boundary (TMA needs 16).
  kv       per layer K and V  [B_max, n_kv, T_max, d] bf16 (head-major: decode streams [T, d] per head).
  act      x[N,H], qkv[N,(n_h+2n_kv)d], q[N,n_h d], attn[N,n_h d], h[N,H], act[N,I]  for N = B*S tokens.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .. import native as nat
from .configs import ShardModelConfig
from .weights import init_state_dict

ALIGN = 128  # elements (256 B)


_GEMV_MAX_ROWS = None


def gemv_max_rows() -> int:
    """Rows per decode step up to which the weight-streaming GEMV path is used; more rows go through the tcgen05 GEMM
    path (one weight pass for all rows; the GEMV kernel needs a second pass above 4 rows and spends CUDA-core FMAs per
    row).  Measured on Qwen2.5-7B (tok/s, GEMV vs GEMM path): 3 rows 802 / 667, 4 rows 851 / 929, 8 rows 880 / 1783 ->
    default 3.  TL_GEMV_MAX_ROWS overrides (1..8)."""
    global _GEMV_MAX_ROWS
    if _GEMV_MAX_ROWS is None:
        import os
        _GEMV_MAX_ROWS = max(1, min(8, int(os.environ.get("TL_GEMV_MAX_ROWS", "3"))))
    return _GEMV_MAX_ROWS


def _rope_inv_freq(cfg: ShardModelConfig) -> torch.Tensor:
    """site-packages/transformers/models/qwen2/modeling_qwen2.py:84-99, computed on the host in fp32 like HF."""
    d = cfg.head_dim
    return 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(torch.float32) / d))


class ShardParams:
    """Flat bf16 parameter arena of one shard, with fused-QKV / interleaved gate-up views."""

    def __init__(self, cfg: ShardModelConfig, layer_ids: Sequence[int], has_embed: bool, has_head: bool,
                 device, with_grad: bool = False):
        self.cfg, self.layer_ids = cfg, list(layer_ids)
        self.has_embed, self.has_head = has_embed, has_head
        self.device = torch.device(device)
        H, I = cfg.hidden, cfg.intermediate
        spec: List[tuple] = []
        for li in self.layer_ids:
            spec += [(f"l{li}.ln1", (H,)), (f"l{li}.wqkv", (cfg.qkv_dim, H))]
            if cfg.qkv_bias:
                spec.append((f"l{li}.bqkv", (cfg.qkv_dim,)))
            if cfg.qk_norm:
                spec += [(f"l{li}.qn", (cfg.head_dim,)), (f"l{li}.kn", (cfg.head_dim,))]
            spec += [(f"l{li}.wo", (H, cfg.q_dim)), (f"l{li}.ln2", (H,)), (f"l{li}.wgu", (2 * I, H)),
                     (f"l{li}.wd", (H, I))]
        if has_embed:
            spec.append(("embed", (cfg.vocab, H)))
        if has_head:
            spec.append(("norm", (H,)))
            if not (cfg.tied and has_embed):
                spec.append(("head", (cfg.vocab, H)))
        self.spec = spec
        self.offsets: Dict[str, tuple] = {}
        off = 0
        for name, shape in spec:
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = (off, n, shape)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.bfloat16, device=self.device)
        self.grad: Optional[torch.Tensor] = torch.zeros_like(self.flat) if with_grad else None
        self.v: Dict[str, torch.Tensor] = {n: self.flat[o:o + k].view(shape) for n, (o, k, shape) in self.offsets.items()}
        self.g: Dict[str, torch.Tensor] = ({n: self.grad[o:o + k].view(shape) for n, (o, k, shape) in self.offsets.items()}
                                           if with_grad else {})
        if has_head and cfg.tied and has_embed:
            self.v["head"] = self.v["embed"]
            if with_grad:
                self.g["head"] = self.g["embed"]

    # ---- HF state dict  <->  fused layout ------------------------------------------------------------
    def load_hf_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Fuse q/k/v, interleave gate/up (the role of worker.py:542-638 ``_load_grouped_layer_weights``)."""
        cfg = self.cfg
        dev = self.device

        def put(name, t):
            self.v[name].copy_(t.to(device=dev, dtype=torch.bfloat16))

        for li in self.layer_ids:
            p = f"model.layers.{li}."
            put(f"l{li}.ln1", sd[p + "input_layernorm.weight"])
            put(f"l{li}.wqkv", torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"],
                                          sd[p + "self_attn.v_proj.weight"]], dim=0))
            if cfg.qkv_bias:
                put(f"l{li}.bqkv", torch.cat([sd[p + "self_attn.q_proj.bias"], sd[p + "self_attn.k_proj.bias"],
                                              sd[p + "self_attn.v_proj.bias"]], dim=0))
            if cfg.qk_norm:
                put(f"l{li}.qn", sd[p + "self_attn.q_norm.weight"])
                put(f"l{li}.kn", sd[p + "self_attn.k_norm.weight"])
            put(f"l{li}.wo", sd[p + "self_attn.o_proj.weight"])
            put(f"l{li}.ln2", sd[p + "post_attention_layernorm.weight"])
            g, u = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
            put(f"l{li}.wgu", torch.stack([g, u], dim=1).reshape(2 * cfg.intermediate, cfg.hidden))
            put(f"l{li}.wd", sd[p + "mlp.down_proj.weight"])
        if self.has_embed:
            put("embed", sd["model.embed_tokens.weight"])
        if self.has_head:
            put("norm", sd["model.norm.weight"])
            if not (cfg.tied and self.has_embed):
                put("head", sd["lm_head.weight"] if "lm_head.weight" in sd else sd["model.embed_tokens.weight"])

    def hf_state_dict(self, grads: bool = False) -> Dict[str, torch.Tensor]:
        """Inverse mapping (the role of ``parameters(distributed=True)``, module.py:577-650)."""
        cfg = self.cfg
        if grads and getattr(self, "grad_settle", None) is not None:
            self.grad_settle()          # matrix gradients are zeroed lazily after zero_grad() (ml/train.py)
        src = self.g if grads else self.v
        out: Dict[str, torch.Tensor] = {}
        for li in self.layer_ids:
            p = f"model.layers.{li}."
            out[p + "input_layernorm.weight"] = src[f"l{li}.ln1"].clone()
            q, k, v = src[f"l{li}.wqkv"].split([cfg.q_dim, cfg.kv_dim, cfg.kv_dim], dim=0)
            out[p + "self_attn.q_proj.weight"], out[p + "self_attn.k_proj.weight"] = q.clone(), k.clone()
            out[p + "self_attn.v_proj.weight"] = v.clone()
            if cfg.qkv_bias:
                bq, bk, bv = src[f"l{li}.bqkv"].split([cfg.q_dim, cfg.kv_dim, cfg.kv_dim], dim=0)
                out[p + "self_attn.q_proj.bias"], out[p + "self_attn.k_proj.bias"] = bq.clone(), bk.clone()
                out[p + "self_attn.v_proj.bias"] = bv.clone()
            if cfg.qk_norm:
                out[p + "self_attn.q_norm.weight"] = src[f"l{li}.qn"].clone()
                out[p + "self_attn.k_norm.weight"] = src[f"l{li}.kn"].clone()
            out[p + "self_attn.o_proj.weight"] = src[f"l{li}.wo"].clone()
            out[p + "post_attention_layernorm.weight"] = src[f"l{li}.ln2"].clone()
            gu = src[f"l{li}.wgu"].view(cfg.intermediate, 2, cfg.hidden)
            out[p + "mlp.gate_proj.weight"], out[p + "mlp.up_proj.weight"] = gu[:, 0].clone(), gu[:, 1].clone()
            out[p + "mlp.down_proj.weight"] = src[f"l{li}.wd"].clone()
        if self.has_embed:
            out["model.embed_tokens.weight"] = src["embed"].clone()
        if self.has_head:
            out["model.norm.weight"] = src["norm"].clone()
            # tied heads appear under both names, like HF's own state_dict()
            out["lm_head.weight"] = out["model.embed_tokens.weight"] if (cfg.tied and self.has_embed) else src["head"].clone()
        return out

    def init_seeded(self, seed: int = 1234):
        """Same values the CPU oracle draws (weights.init_state_dict), materialised layer by layer."""
        cfg = self.cfg
        for li in self.layer_ids:
            sd = init_state_dict(cfg, seed, torch.bfloat16, "cpu", layers=[li], with_embed=False, with_head=False)
            sub = ShardParams.__new__(ShardParams)
            sub.__dict__.update(self.__dict__)
            sub.layer_ids, sub.has_embed, sub.has_head = [li], False, False
            sub.load_hf_state_dict(sd)
        sd = init_state_dict(cfg, seed, torch.bfloat16, "cpu", layers=[], with_embed=self.has_embed or (
            self.has_head and cfg.tied), with_head=self.has_head)
        sub = ShardParams.__new__(ShardParams)
        sub.__dict__.update(self.__dict__)
        sub.layer_ids = []
        sub.load_hf_state_dict(sd)

    def init_on_device(self, seed: int = 1234, std: float = 0.02):
        """Random init drawn directly on the GPU (benchmarks at 7B/8B scale; not comparable with the oracle)."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        self.flat.normal_(0.0, std, generator=g)
        for name in self.v:
            if name.endswith(("ln1", "ln2", "qn", "kn")) or name == "norm":
                self.v[name].fill_(1.0)


@dataclass
class ShardBuffers:
    """Activation workspaces for up to ``n_max`` tokens in flight."""
    x: torch.Tensor
    h: torch.Tensor
    qkv: torch.Tensor
    q: torch.Tensor
    attn: torch.Tensor
    act: torch.Tensor


class CudaLayerGroup:
    """B200 shard operator (LayerGroupModule mirror, injector.py:154-281)."""

    def __init__(self, cfg: ShardModelConfig, params: ShardParams, max_batch: int, max_seq: int,
                 max_tokens: Optional[int] = None):
        nat.require_device()
        self.cfg, self.p = cfg, params
        self.layer_ids = params.layer_ids
        self.num_layers = len(self.layer_ids)             # worker.py:332-335 dispatches on this attribute
        self.input_vars = ["hidden_states", "past_len", "use_cache"]
        self.output_vars = ["hidden_states"]
        self.device = params.device
        self.B_max, self.T_max = max_batch, max_seq
        dev, bf = self.device, torch.bfloat16
        self.kc = [torch.zeros(max_batch, cfg.n_kv_heads, max_seq, cfg.head_dim, dtype=bf, device=dev)
                   for _ in self.layer_ids]
        self.vc = [torch.zeros_like(k) for k in self.kc]
        self.cos, self.sin = nat.rope_table(_rope_inv_freq(cfg).to(dev), max_seq)
        # what the launch after this shard's last decode GEMV streams (a prefetch hint only, see _layer_decode)
        self.weights_after_last_layer = (params.v.get("head") if "head" in params.v else
                                         params.v.get(f"l{self.layer_ids[0]}.wqkv") if len(self.layer_ids) else None)
        self.pos_dev = torch.zeros(1, dtype=torch.int32, device=dev)      # next write position in the cache
        self.kvlen_dev = torch.zeros(1, dtype=torch.int32, device=dev)    # valid keys for the decode kernel
        self.n_max = max_tokens or max_batch * max_seq
        self._alloc_bufs(min(self.n_max, 8))
        self.dbufs = self._make_bufs(max_batch)       # decode-time buffers: fixed addresses (captured graphs, job lists)
        # split-K workspace for batched decode (rows above the GEMV threshold, <= 128): the qkv / o / down Linears have
        # too few output tiles to occupy every SM
        self.gemm_ws = (torch.empty(nat.gemm_splitk_ws(min(max_batch, 128), max(cfg.qkv_dim, cfg.hidden)), dtype=torch.uint8,
                                    device=dev) if max_batch > 1 else None)
        self.dec_ws = torch.empty(max(nat.attn_decode_ws(max_batch, cfg.n_heads, cfg.head_dim, max_seq), 16),
                                  dtype=torch.uint8, device=dev)
        self.scale = cfg.head_dim ** -0.5
        self.allow_chain = True              # DistributedModel clears it when NCCL kernels share the device during decode
        self.chain_sync: Optional[torch.Tensor] = None
        self.chain_attn_ws: Optional[torch.Tensor] = None
        self._chains: Dict[tuple, list] = {}

    def _make_bufs(self, n: int) -> ShardBuffers:
        cfg, dev, bf = self.cfg, self.device, torch.bfloat16
        return ShardBuffers(
            x=torch.empty(n, cfg.hidden, dtype=bf, device=dev), h=torch.empty(n, cfg.hidden, dtype=bf, device=dev),
            qkv=torch.empty(n, cfg.qkv_dim, dtype=bf, device=dev), q=torch.empty(n, cfg.q_dim, dtype=bf, device=dev),
            attn=torch.empty(n, cfg.q_dim, dtype=bf, device=dev),
            act=torch.empty(n, cfg.intermediate, dtype=bf, device=dev))

    def _alloc_bufs(self, n: int):
        self.bufs = self._make_bufs(n)
        self.n_alloc = n

    def _dbufs(self, n: int) -> ShardBuffers:
        b = self.dbufs
        return ShardBuffers(b.x[:n], b.h[:n], b.qkv[:n], b.q[:n], b.attn[:n], b.act[:n])

    def _bufs(self, n: int) -> ShardBuffers:
        if n > self.n_alloc:
            self._alloc_bufs(n)
        b = self.bufs
        return ShardBuffers(b.x[:n], b.h[:n], b.qkv[:n], b.q[:n], b.attn[:n], b.act[:n])

    # ------------------------------------------------------------------------------------------ cache control
    def reset_cache(self, past_len: int = 0):
        self.pos_dev.fill_(past_len)
        self.kvlen_dev.fill_(past_len)

    # ------------------------------------------------------------------------------------------ layer bodies
    def _layer_prefill(self, j: int, x: torch.Tensor, B: int, S: int, past_len: int, w: ShardBuffers):
        """site-packages/transformers/models/qwen2/modeling_qwen2.py:280-310 for N = B*S tokens (GEMM path)."""
        cfg, v, li = self.cfg, self.p.v, self.layer_ids[j]
        nat.rmsnorm_fwd(x, v[f"l{li}.ln1"], cfg.rms_eps, out=w.h)
        nat.gemm(w.h, v[f"l{li}.wqkv"], out=w.qkv, bias=v.get(f"l{li}.bqkv"))
        nat.rope_kv_fwd(w.qkv, w.q, self.kc[j], self.vc[j], self.pos_dev, self.cos, self.sin, v.get(f"l{li}.qn"),
                        v.get(f"l{li}.kn"), cfg.rms_eps, S, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim)
        nat.attn_prefill_fwd(w.q, self.kc[j], self.vc[j], w.attn, None, B, S, past_len, cfg.n_heads, cfg.n_kv_heads,
                             cfg.head_dim, self.scale)
        nat.gemm(w.attn, v[f"l{li}.wo"], out=x, residual=x)
        nat.rmsnorm_fwd(x, v[f"l{li}.ln2"], cfg.rms_eps, out=w.h)
        nat.gemm(w.h, v[f"l{li}.wgu"], out=w.act, flags=nat.EPI_SWIGLU)
        nat.gemm(w.act, v[f"l{li}.wd"], out=x, residual=x)

    FUSED_DECODE_MAX_T = 2048

    def _decode_attention(self, j: int, li: int, B: int, w: ShardBuffers):
        """w.qkv (post-bias) -> w.attn for one new token per row.  Short caches: one fused launch (RoPE + append +
        attention); long caches: RoPE/append, split-KV partials, reduce (three launches, parallel over the KV length)."""
        cfg, v = self.cfg, self.p.v
        if self.T_max <= self.FUSED_DECODE_MAX_T and (cfg.n_heads // cfg.n_kv_heads) <= 8:
            nat.attn_decode_fused(w.qkv, self.kc[j], self.vc[j], w.attn, self.pos_dev, self.cos, self.sin,
                                  v.get(f"l{li}.qn"), v.get(f"l{li}.kn"), cfg.rms_eps, B, cfg.n_heads, cfg.n_kv_heads,
                                  cfg.head_dim, self.scale)
            return
        nat.rope_kv_fwd(w.qkv, w.q, self.kc[j], self.vc[j], self.pos_dev, self.cos, self.sin, v.get(f"l{li}.qn"),
                        v.get(f"l{li}.kn"), cfg.rms_eps, 1, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim)
        nat.attn_decode_fwd(w.q, self.kc[j], self.vc[j], w.attn, self.kvlen_dev, self.dec_ws, B, cfg.n_heads,
                            cfg.n_kv_heads, cfg.head_dim, self.scale)

    def _layer_decode(self, j: int, x: torch.Tensor, B: int, w: ShardBuffers, out: Optional[torch.Tensor] = None):
        """Same layer for B <= 8 single-token rows: weight-streaming GEMVs with the norms fused as prologues."""
        cfg, v, li = self.cfg, self.p.v, self.layer_ids[j]
        # every GEMV names the weights of the launch after it: it queues L2 prefetches behind its own loads, so HBM keeps
        # streaming through the launch boundary (and through the attention kernel) instead of idling there
        if j + 1 < self.num_layers:
            after = v[f"l{self.layer_ids[j + 1]}.wqkv"]
        else:
            after = self.weights_after_last_layer            # lm_head on the last stage, else layer 0 for the next slot
        nat.gemv(x, v[f"l{li}.wqkv"], out=w.qkv, bias=v.get(f"l{li}.bqkv"), norm_w=v[f"l{li}.ln1"], eps=cfg.rms_eps,
                 next_w=v[f"l{li}.wo"])
        self._decode_attention(j, li, B, w)
        nat.gemv(w.attn, v[f"l{li}.wo"], out=x, residual=x, next_w=v[f"l{li}.wgu"])
        nat.gemv(x, v[f"l{li}.wgu"], out=w.act, norm_w=v[f"l{li}.ln2"], eps=cfg.rms_eps, flags=nat.EPI_SWIGLU,
                 next_w=v[f"l{li}.wd"])
        nat.gemv(w.act, v[f"l{li}.wd"], out=x if out is None else out, residual=x, next_w=after)

    def _layer_decode_batched(self, j: int, x: torch.Tensor, B: int, w: ShardBuffers, out: Optional[torch.Tensor] = None):
        """More single-token rows than the GEMV path takes: tcgen05 GEMMs in the weight-streaming regime (split along K
        where a Linear has too few output tiles to occupy every SM) + decode attention.  The RMSNorm after each
        residual Linear rides in that Linear's split-K reduce pass, so layer j > 0 finds its normalised input in w.h."""
        cfg, v, li = self.cfg, self.p.v, self.layer_ids[j]
        ws = self.gemm_ws if B <= 128 else None
        if j == 0:
            nat.rmsnorm_fwd(x, v[f"l{li}.ln1"], cfg.rms_eps, out=w.h)
        nat.gemm(w.h, v[f"l{li}.wqkv"], out=w.qkv, bias=v.get(f"l{li}.bqkv"), ws=ws)
        self._decode_attention(j, li, B, w)
        nat.gemm(w.attn, v[f"l{li}.wo"], out=x, residual=x, ws=ws, norm_w=v[f"l{li}.ln2"], eps=cfg.rms_eps, h_out=w.h)
        nat.gemm(w.h, v[f"l{li}.wgu"], out=w.act, flags=nat.EPI_SWIGLU, ws=ws)
        if j + 1 < self.num_layers:
            nat.gemm(w.act, v[f"l{li}.wd"], out=x, residual=x, ws=ws, norm_w=v[f"l{self.layer_ids[j + 1]}.ln1"],
                     eps=cfg.rms_eps, h_out=w.h)
        else:
            nat.gemm(w.act, v[f"l{li}.wd"], out=x if out is None else out, residual=x, ws=ws)

    def prefill(self, hidden: torch.Tensor, past_len: int = 0) -> torch.Tensor:
        """hidden [B,S,H] -> [B,S,H]; appends S positions to the KV cache starting at ``past_len``."""
        B, S, H = hidden.shape
        if B > self.B_max or past_len + S > self.T_max:
            raise ValueError(f"shard sized for B<={self.B_max}, T<={self.T_max}; got B={B}, T={past_len + S}")
        N = B * S
        w = self._bufs(N)
        w.x.copy_(hidden.reshape(N, H))
        self.pos_dev.fill_(past_len)
        for j in range(self.num_layers):
            self._layer_prefill(j, w.x, B, S, past_len, w)
        self.pos_dev.fill_(past_len + S)
        self.kvlen_dev.fill_(past_len + S)
        return w.x.view(B, S, H)

    def decode_step_inplace(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, advance: bool = True):
        """x [B,H] updated in place through this shard's layers; one new token per row at position ``pos_dev``.
        Graph-capturable: the write position and the KV length live in device memory and are advanced by
        kernels inside the same launch sequence (kv_len += 1 before the layers, pos += 1 after).
        ``out``: where the LAST layer's down projection stores the shard's output rows instead of ``x`` — the next
        stage's peer-mapped input buffer (p2p/peer.py), so the hop rides on that kernel's own stores.
        ``advance=False``: the caller advances ``kvlen_dev`` before and ``pos_dev`` after (the mailbox kernels do)."""
        B = x.shape[0]
        w = self._dbufs(B)
        if advance:
            nat.advance_pos(self.kvlen_dev, None, 1)
        if self.chain_ok(B) or self.dq_ok(B):
            (self._decode_step_chained if self.chain_ok(B) else self._decode_step_dq)(x, B, out)
            if advance:
                nat.advance_pos(self.pos_dev, None, 1)
            return
        for j in range(self.num_layers):
            o = out if j == self.num_layers - 1 else None
            if B <= gemv_max_rows():
                self._layer_decode(j, x, B, w, o)
            else:
                self._layer_decode_batched(j, x, B, w, o)
        if advance:
            nat.advance_pos(self.pos_dev, None, 1)

    # ------------------------------------------------------------------------------------------ chained decode step
    def chain_ok(self, B: int) -> bool:
        """Rows / shapes the persistent chain kernel (csrc/decode_chain.cu) takes.  Opt-in (TL_DECODE_IMPL=chain): measured
        on Qwen2.5-7B it streams every Linear at the HBM rate (the GEMV phases of a layer sum to 69-75 us against a 72 us
        bound) but pays 6-7 us per software dependency (release/acquire counter + restaging the input vector) where a
        programmatic-dependent-launch boundary costs ~4 us: 323 tok/s against 355+ for the per-kernel sequence
        (profiles/r02_decode_chain_timeline.txt)."""
        import os
        cfg = self.cfg
        return (os.environ.get("TL_DECODE_IMPL", "kernels") == "chain" and self.allow_chain and self.num_layers > 0
                and B <= min(4, gemv_max_rows()) and cfg.n_kv_heads * B <= 60 and cfg.n_heads // cfg.n_kv_heads <= 8
                and cfg.head_dim in (64, 128))

    def _chain_group(self) -> int:
        """Decoder layers per chain launch (TL_CHAIN_LAYERS, default 1; 5 jobs per layer, at most 3 layers)."""
        import os
        return max(1, min(3, int(os.environ.get("TL_CHAIN_LAYERS", "1"))))

    def _decode_chains(self, x: torch.Tensor, B: int, out: Optional[torch.Tensor]):
        """The launch list of one decode step of this shard: qkv of the first layer as a stand-alone GEMV, then one
        persistent launch per group of layers [ATTN, o, gate/up, down, qkv of the next layer]."""
        key = (B, x.data_ptr(), 0 if out is None else out.data_ptr())
        if key in self._chains:
            return self._chains[key]
        cfg, v = self.cfg, self.p.v
        w = self._dbufs(B)
        J = nat.make_job
        if self.chain_sync is None:
            self.chain_sync = torch.zeros(self.num_layers + 1, nat.CHAIN_SYNC_BYTES // 4, dtype=torch.int32, device=self.device)
            self.chain_attn_ws = torch.empty(nat.decode_chain_ws(min(self.B_max, 4), cfg.n_heads, cfg.n_kv_heads, cfg.head_dim),
                                             dtype=torch.uint8, device=self.device)
        per = self._chain_group()
        launches = []
        for g0 in range(0, self.num_layers, per):
            jobs = []
            g1 = min(self.num_layers, g0 + per)
            for j in range(g0, g1):
                li = self.layer_ids[j]
                jobs.append(J(nat.JOB_ATTN, x=w.qkv, y=w.attn, k_cache=self.kc[j], v_cache=self.vc[j], pos_dev=self.pos_dev,
                              cos_tab=self.cos, sin_tab=self.sin, q_norm_w=v.get(f"l{li}.qn"), k_norm_w=v.get(f"l{li}.kn"),
                              n_h=cfg.n_heads, n_kv=cfg.n_kv_heads, d=cfg.head_dim, T_max=self.T_max, scale=self.scale,
                              eps=cfg.rms_eps))
                jobs.append(J(nat.JOB_GEMV, N=cfg.hidden, K=cfg.q_dim, flags=nat.EPI_RESIDUAL, W=v[f"l{li}.wo"], x=w.attn, y=x,
                              residual=x))
                jobs.append(J(nat.JOB_GEMV, N=2 * cfg.intermediate, K=cfg.hidden, flags=nat.EPI_SWIGLU, W=v[f"l{li}.wgu"], x=x,
                              y=w.act, norm_w=v[f"l{li}.ln2"], eps=cfg.rms_eps))
                last = j == self.num_layers - 1
                jobs.append(J(nat.JOB_GEMV, N=cfg.hidden, K=cfg.intermediate, flags=nat.EPI_RESIDUAL, W=v[f"l{li}.wd"], x=w.act,
                              y=(out if (last and out is not None) else x), residual=x))
                if not last:
                    ln = self.layer_ids[j + 1]
                    bq = v.get(f"l{ln}.bqkv")
                    jobs.append(J(nat.JOB_GEMV, N=cfg.qkv_dim, K=cfg.hidden, flags=nat.EPI_BIAS if bq is not None else 0,
                                  W=v[f"l{ln}.wqkv"], x=x, y=w.qkv, bias=bq, norm_w=v[f"l{ln}.ln1"], eps=cfg.rms_eps))
            # what the launch after this one streams first: the next group's o-projection, or whatever follows the shard
            nxt = v[f"l{self.layer_ids[g1]}.wo"] if g1 < self.num_layers else self.weights_after_last_layer
            launches.append(nat.DecodeChain(jobs, B, self.chain_sync[g0 // per], self.chain_attn_ws, nxt))
        self._chains[key] = launches
        return launches

    def _decode_step_chained(self, x: torch.Tensor, B: int, out: Optional[torch.Tensor]):
        cfg, v = self.cfg, self.p.v
        w = self._dbufs(B)
        l0 = self.layer_ids[0]
        nat.gemv(x, v[f"l{l0}.wqkv"], out=w.qkv, bias=v.get(f"l{l0}.bqkv"), norm_w=v[f"l{l0}.ln1"], eps=cfg.rms_eps,
                 next_w=v[f"l{l0}.wo"])
        for ch in self._decode_chains(x, B, out):
            ch.launch()

    def dq_ok(self, B: int) -> bool:
        """TL_DECODE_IMPL=dq: the per-kernel sequence, except that the down projection of layer j and the qkv projection
        of layer j+1 run as ONE two-job chain launch (one software dependency instead of a launch boundary between a
        long and a short weight stream)."""
        import os
        return os.environ.get("TL_DECODE_IMPL", "kernels") == "dq" and self.allow_chain and self.num_layers > 1 and B <= min(4, gemv_max_rows())

    def _decode_step_dq(self, x: torch.Tensor, B: int, out: Optional[torch.Tensor]):
        cfg, v = self.cfg, self.p.v
        w = self._dbufs(B)
        key = ("dq", B, x.data_ptr(), 0 if out is None else out.data_ptr())
        if key not in self._chains:
            J = nat.make_job
            if self.chain_sync is None:
                self.chain_sync = torch.zeros(self.num_layers + 1, nat.CHAIN_SYNC_BYTES // 4, dtype=torch.int32, device=self.device)
                self.chain_attn_ws = torch.empty(nat.decode_chain_ws(min(self.B_max, 4), cfg.n_heads, cfg.n_kv_heads, cfg.head_dim),
                                                 dtype=torch.uint8, device=self.device)
            launches = []
            for j in range(self.num_layers - 1):
                li, ln = self.layer_ids[j], self.layer_ids[j + 1]
                bq = v.get(f"l{ln}.bqkv")
                jobs = [J(nat.JOB_GEMV, N=cfg.hidden, K=cfg.intermediate, flags=nat.EPI_RESIDUAL, W=v[f"l{li}.wd"], x=w.act, y=x, residual=x),
                        J(nat.JOB_GEMV, N=cfg.qkv_dim, K=cfg.hidden, flags=nat.EPI_BIAS if bq is not None else 0, W=v[f"l{ln}.wqkv"], x=x,
                          y=w.qkv, bias=bq, norm_w=v[f"l{ln}.ln1"], eps=cfg.rms_eps)]
                launches.append(nat.DecodeChain(jobs, B, self.chain_sync[j], self.chain_attn_ws, v[f"l{ln}.wo"]))
            self._chains[key] = launches
        chains = self._chains[key]
        l0 = self.layer_ids[0]
        nat.gemv(x, v[f"l{l0}.wqkv"], out=w.qkv, bias=v.get(f"l{l0}.bqkv"), norm_w=v[f"l{l0}.ln1"], eps=cfg.rms_eps, next_w=v[f"l{l0}.wo"])
        for j, li in enumerate(self.layer_ids):
            self._decode_attention(j, li, B, w)
            nat.gemv(w.attn, v[f"l{li}.wo"], out=x, residual=x, next_w=v[f"l{li}.wgu"])
            nat.gemv(x, v[f"l{li}.wgu"], out=w.act, norm_w=v[f"l{li}.ln2"], eps=cfg.rms_eps, flags=nat.EPI_SWIGLU, next_w=v[f"l{li}.wd"])
            if j + 1 < self.num_layers:
                chains[j].launch()
            else:
                nat.gemv(w.act, v[f"l{li}.wd"], out=x if out is None else out, residual=x, next_w=self.weights_after_last_layer)

    def n_chain_launches(self) -> int:
        per = self._chain_group()
        return 1 + (self.num_layers + per - 1) // per

    def check(self):
        """Raise if a chain launch gave up on a dependency wait (error word of its sync slot)."""
        if self.chain_sync is not None and int(self.chain_sync[:, 2].max().item()):
            self.chain_sync[:, :4].zero_()
            raise nat.NativeError("decode chain kernel timed out waiting for a dependency (co-residency of its CTAs lost?)")

    # ------------------------------------------------------------------------------------------ reference-shaped API
    def forward(self, **kwargs) -> dict:
        """``LayerGroupModule.forward`` contract (injector.py:252-260): returns kwargs ∪ outputs."""
        hs = kwargs["hidden_states"]
        past_len = int(kwargs.get("past_len", 0) or 0)
        out = dict(kwargs)
        out["hidden_states"] = self.prefill(hs, past_len).clone()
        return out

    __call__ = forward
