"""CPU oracle for the tensorlink shard-executor hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product (``tensorlink_b200``) never does: it fails loudly without its CUDA library.

What is restated, and from where (paths relative to /root/reference unless they begin with
``site-packages/``, which is the third-party ``transformers`` the reference ``exec``s verbatim):

* a shard = a contiguous layer range applied in order to ``hidden_states`` with shared
  ``position_embeddings``; shards compose sequentially; the wire between them is lossless
  (tensorlink/ml/injector.py:154-281 ``LayerGroupModule``; :508-556 ``_generate_worker_calls``;
  tensorlink/ml/utils.py:569-660 codec)                                   -> ``shard_forward``
* embed / final norm / lm_head live outside the shards, tied head shares the embedding
  (tensorlink/ml/module.py:1023-1056, :1218-1265)                         -> ``model_forward``
* per-layer math (site-packages/transformers/models/qwen2/modeling_qwen2.py):
  RMSNorm :258-263, rotary tables :102-113, rotate_half/apply :116-146, attention :206-245 with
  eager softmax :161-184, MLP :46-48, decoder layer :280-310; Qwen3 q/k-norm
  (…/qwen3/modeling_qwen3.py:248-264)
* backward = autograd over exactly these ops, grads of shard inputs returned upstream, missing
  grads -> zeros (tensorlink/ml/worker.py:233-295); optimizer = plain ``optimizer.step()``
  (tensorlink/ml/worker.py:1309-1327)
* generate = greedy argmax of the last position with a KV cache (tensorlink/ml/module.py:763-769
  delegates to HF ``generate``; ``do_sample=False``)

PARITY PINNING: the reference's own tests hold no numeric vector for this path
(tests/test_distributed_model.py:27-77 assert nothing) -> "parity unpinned" at the reference
boundary.  This oracle is instead pinned against (a) installed HF ``Qwen2ForCausalLM`` /
``Qwen3ForCausalLM`` on CPU, bit-exact with ``attn_mode='eager'`` (tests/test_oracle_vs_hf.py), and
(b) the reference's own ``LayerGroupModule`` + wire codec imported through ``oracle/ref_shim.py``
(``oracle/gen_golden.py`` -> tests/golden/ref_layergroup_*.pt).

``attn_mode``:
  'eager'      bit-identical to HF ``eager_attention_forward`` (scores rounded to the activation
               dtype before the fp32 softmax).
  'sdpa_math'  the SDPA contract HF uses by default (``_attn_implementation='sdpa'``): fp32
               scores and softmax, probabilities cast to the activation dtype for P@V, fp32
               accumulate.  This is what the CUDA flash kernels are compared with.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from tensorlink_b200.ml.configs import ShardModelConfig


# --------------------------------------------------------------------------- per-op restatements
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """modeling_qwen2.py:258-263 — fp32 normalise, cast to input dtype, THEN multiply by gain."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rope_inv_freq(cfg: ShardModelConfig) -> torch.Tensor:
    """modeling_qwen2.py:84-99 (default rope init)."""
    d = cfg.head_dim
    return 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(torch.float32) / d))


def rope_tables(cfg: ShardModelConfig, position_ids: torch.Tensor, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """modeling_qwen2.py:102-113 — fp32 outer product, cos/sin, cast to activation dtype. [B,S,d]."""
    inv = rope_inv_freq(cfg)[None, :, None].expand(position_ids.shape[0], -1, 1)
    pos = position_ids[:, None, :].to(torch.float32)
    freqs = (inv @ pos).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin):
    """modeling_qwen2.py:124-146 — q,k are [B,heads,S,d]; cos/sin [B,S,d]."""
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def repeat_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    if n_rep == 1:
        return x
    b, h, t, d = x.shape
    return x[:, :, None].expand(b, h, n_rep, t, d).reshape(b, h * n_rep, t, d)


def causal_mask(S: int, T: int, dtype) -> torch.Tensor:
    """Additive mask for queries at absolute positions T-S..T-1 over keys 0..T-1."""
    qpos = torch.arange(T - S, T)[:, None]
    kpos = torch.arange(T)[None, :]
    m = torch.zeros(S, T, dtype=dtype)
    m.masked_fill_(kpos > qpos, torch.finfo(dtype).min)
    return m[None, None]


def attention_eager(q, k, v, scaling: float, n_rep: int) -> torch.Tensor:
    """modeling_qwen2.py:161-184.  q [B,n_h,S,d]; k,v [B,n_kv,T,d] -> [B,S,n_h*d]."""
    k = repeat_kv(k, n_rep)
    v = repeat_kv(v, n_rep)
    S, T = q.shape[2], k.shape[2]
    w = torch.matmul(q, k.transpose(2, 3)) * scaling
    w = w + causal_mask(S, T, q.dtype)
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)
    return o.transpose(1, 2).contiguous().reshape(q.shape[0], S, -1)


def attention_sdpa_math(q, k, v, scaling: float, n_rep: int) -> torch.Tensor:
    """SDPA contract: fp32 scores/softmax from exact bf16 products, P cast to dtype, fp32 PV."""
    dt = q.dtype
    k = repeat_kv(k, n_rep).to(torch.float32)
    v = repeat_kv(v, n_rep)
    S, T = q.shape[2], k.shape[2]
    s = torch.matmul(q.to(torch.float32), k.transpose(2, 3)) * scaling
    s = s + causal_mask(S, T, torch.float32)
    p = F.softmax(s, dim=-1)
    o = torch.matmul(p.to(dt).to(torch.float32), v.to(torch.float32)).to(dt)
    return o.transpose(1, 2).contiguous().reshape(q.shape[0], S, -1)


def swiglu_mlp(x, wg, wu, wd):
    """modeling_qwen2.py:46-48."""
    return F.linear(F.silu(F.linear(x, wg)) * F.linear(x, wu), wd)


# --------------------------------------------------------------------------- containers
@dataclass
class LayerWeights:
    ln1: torch.Tensor
    wq: torch.Tensor
    wk: torch.Tensor
    wv: torch.Tensor
    wo: torch.Tensor
    ln2: torch.Tensor
    wg: torch.Tensor
    wu: torch.Tensor
    wd: torch.Tensor
    bq: Optional[torch.Tensor] = None
    bk: Optional[torch.Tensor] = None
    bv: Optional[torch.Tensor] = None
    qn: Optional[torch.Tensor] = None
    kn: Optional[torch.Tensor] = None

    @staticmethod
    def from_state_dict(sd: Dict[str, torch.Tensor], li: int) -> "LayerWeights":
        p = f"model.layers.{li}."
        g = lambda n: sd.get(p + n)
        return LayerWeights(
            ln1=g("input_layernorm.weight"), wq=g("self_attn.q_proj.weight"),
            wk=g("self_attn.k_proj.weight"), wv=g("self_attn.v_proj.weight"),
            wo=g("self_attn.o_proj.weight"), ln2=g("post_attention_layernorm.weight"),
            wg=g("mlp.gate_proj.weight"), wu=g("mlp.up_proj.weight"), wd=g("mlp.down_proj.weight"),
            bq=g("self_attn.q_proj.bias"), bk=g("self_attn.k_proj.bias"), bv=g("self_attn.v_proj.bias"),
            qn=g("self_attn.q_norm.weight"), kn=g("self_attn.k_norm.weight"))

    def tensors(self) -> List[torch.Tensor]:
        return [t for t in (self.ln1, self.wq, self.wk, self.wv, self.wo, self.ln2, self.wg, self.wu,
                            self.wd, self.bq, self.bk, self.bv, self.qn, self.kn) if t is not None]


@dataclass
class KVCache:
    """Per-layer growing K/V, [B,n_kv,T,d] (the role of HF DynamicCache, utils.py:599-605)."""
    k: Dict[int, torch.Tensor] = field(default_factory=dict)
    v: Dict[int, torch.Tensor] = field(default_factory=dict)

    def update(self, li: int, k: torch.Tensor, v: torch.Tensor):
        if li in self.k:
            self.k[li] = torch.cat([self.k[li], k], dim=2)
            self.v[li] = torch.cat([self.v[li], v], dim=2)
        else:
            self.k[li], self.v[li] = k, v
        return self.k[li], self.v[li]

    def length(self) -> int:
        return next(iter(self.k.values())).shape[2] if self.k else 0


# --------------------------------------------------------------------------- layer / shard / model
def decoder_layer(cfg: ShardModelConfig, w: LayerWeights, x: torch.Tensor, cos, sin,
                  attn_mode: str = "sdpa_math", cache: Optional[KVCache] = None, li: int = 0,
                  trace: Optional[dict] = None) -> torch.Tensor:
    """modeling_qwen2.py:280-310 (+ Qwen3 q/k norm)."""
    B, S, _ = x.shape
    d = cfg.head_dim
    residual = x
    h = rmsnorm(x, w.ln1, cfg.rms_eps)
    if trace is not None:
        trace["ln1"] = h
    q = F.linear(h, w.wq, w.bq).view(B, S, -1, d)
    k = F.linear(h, w.wk, w.bk).view(B, S, -1, d)
    v = F.linear(h, w.wv, w.bv).view(B, S, -1, d)
    if cfg.qk_norm:
        q = rmsnorm(q, w.qn, cfg.rms_eps)
        k = rmsnorm(k, w.kn, cfg.rms_eps)
    q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    if trace is not None:
        trace["q_rope"], trace["k_rope"], trace["v"] = q, k, v
    if cache is not None:
        k, v = cache.update(li, k, v)
    fn = attention_eager if attn_mode == "eager" else attention_sdpa_math
    a = fn(q, k, v, d ** -0.5, cfg.n_heads // cfg.n_kv_heads)
    if trace is not None:
        trace["attn"] = a
    x = residual + F.linear(a, w.wo)
    if trace is not None:
        trace["post_attn"] = x
    residual = x
    h = rmsnorm(x, w.ln2, cfg.rms_eps)
    x = residual + swiglu_mlp(h, w.wg, w.wu, w.wd)
    return x


def shard_forward(cfg: ShardModelConfig, layers: Sequence[LayerWeights], layer_ids: Sequence[int],
                  hidden_states: torch.Tensor, cos, sin, attn_mode: str = "sdpa_math",
                  cache: Optional[KVCache] = None) -> torch.Tensor:
    """One ``offloaded_group`` shard: the loop body over its layers (injector.py:236-281)."""
    for w, li in zip(layers, layer_ids):
        hidden_states = decoder_layer(cfg, w, hidden_states, cos, sin, attn_mode, cache, li)
    return hidden_states


def wire_hop(t: torch.Tensor) -> torch.Tensor:
    """The inter-shard hop.  The reference codec is byte-exact for tensor data (utils.py:569-660,
    verified in oracle/gen_golden.py), so the oracle hop is a detached copy."""
    return t.detach().clone()


def split_layers(n_layers: int, n_shards: int) -> List[range]:
    """Even contiguous split; the first ``n_layers % n_shards`` shards take one extra layer."""
    base, rem = divmod(n_layers, n_shards)
    out, a = [], 0
    for i in range(n_shards):
        b = a + base + (1 if i < rem else 0)
        out.append(range(a, b))
        a = b
    return out


class OracleModel:
    """Whole-model composition: embed -> shards -> final norm -> lm_head."""

    def __init__(self, cfg: ShardModelConfig, sd: Dict[str, torch.Tensor], attn_mode: str = "sdpa_math"):
        self.cfg, self.sd, self.attn_mode = cfg, sd, attn_mode
        self.layers = [LayerWeights.from_state_dict(sd, i) for i in range(cfg.n_layers)]
        self.embed = sd["model.embed_tokens.weight"]
        self.norm = sd["model.norm.weight"]
        self.head = sd["lm_head.weight"]

    def hidden(self, input_ids: torch.Tensor, n_shards: int = 1, cache: Optional[KVCache] = None,
               past_len: int = 0, per_layer: Optional[list] = None) -> torch.Tensor:
        cfg = self.cfg
        B, S = input_ids.shape
        x = F.embedding(input_ids, self.embed)
        pos = torch.arange(past_len, past_len + S)[None].expand(B, -1)
        cos, sin = rope_tables(cfg, pos, x.dtype)
        for r in split_layers(cfg.n_layers, n_shards):
            if per_layer is None:
                x = shard_forward(cfg, [self.layers[i] for i in r], list(r), x, cos, sin,
                                  self.attn_mode, cache)
            else:
                for i in r:
                    x = decoder_layer(cfg, self.layers[i], x, cos, sin, self.attn_mode, cache, i)
                    per_layer.append(x)
            if n_shards > 1:
                x = wire_hop(x) if not x.requires_grad else x
        return x

    def logits(self, input_ids, n_shards: int = 1, cache=None, past_len: int = 0,
               last_only: bool = False) -> torch.Tensor:
        x = self.hidden(input_ids, n_shards, cache, past_len)
        if last_only:
            x = x[:, -1:, :]
        x = rmsnorm(x, self.norm, self.cfg.rms_eps)
        return F.linear(x, self.head)

    def loss(self, input_ids, labels, n_shards: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """HF ForCausalLMLoss: fp32 logits, labels shifted left, mean CE, ignore_index -100."""
        logits = self.logits(input_ids, n_shards)
        lf = logits.to(torch.float32)
        shift = F.pad(labels, (0, 1), value=-100)[:, 1:]
        loss = F.cross_entropy(lf.reshape(-1, lf.shape[-1]), shift.reshape(-1), ignore_index=-100)
        return loss, logits

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int, n_shards: int = 1,
                 return_margins: bool = False):
        """Greedy decode with KV cache, EOS disabled.  Returns [B, S+new] int64."""
        cache = KVCache()
        ids = input_ids
        margins = []
        logits = self.logits(ids, n_shards, cache, 0, last_only=True)
        for step in range(max_new_tokens):
            lf = logits[:, -1, :].to(torch.float32)
            nxt = lf.argmax(-1, keepdim=True)
            if return_margins:
                top2 = lf.topk(2, dim=-1).values
                margins.append(top2[:, 0] - top2[:, 1])
            ids = torch.cat([ids, nxt], dim=1)
            if step + 1 < max_new_tokens:
                logits = self.logits(nxt, n_shards, cache, ids.shape[1] - 1, last_only=True)
        if return_margins:
            return ids, torch.stack(margins, dim=1)
        return ids


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().to(torch.float64).flatten().cpu()
    b = b.detach().to(torch.float64).flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
