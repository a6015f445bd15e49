"""Training through the pipeline stages: forward with saved activations, hand-written backward, fused AdamW.

Reference semantics being replaced (/root/reference/tensorlink/ml):
  * ``DistributedModel.forward`` splits the batch into ``n_pipelines`` micro-batches (module.py:374-399) and wraps the
    output in ``CustomAutogradRouter`` so that ``loss.backward()`` reaches ``DistributedModel.backward`` (:126-144,
    :414-437), which walks the recorded shard boundaries LIFO and RPCs each gradient to its worker (:439-524);
  * the worker keeps ``{"inputs","output"}`` per micro-batch and runs ``assoc_output.backward(grad)`` over HF's
    autograd graph, returning the gradient of the shard input (worker.py:233-295, zeros where a grad is missing);
  * ``optimizer.step()`` / ``zero_grad()`` fan out to every worker (optim.py:131-187, worker.py:1309-1327).
Here each rank owns its stage's flat parameter and gradient arenas; the backward of a decoder layer is eight
tcgen05 GEMMs (dgrad + wgrad, MN-major operands, gradient accumulation in the epilogue) plus the attention /
norm / RoPE / SwiGLU backward kernels; gradients of ``hidden_states`` hop rank i+1 -> i over NVLink.  The loss and
its gradient are produced on the last stage by a fused lm_head + cross-entropy pass over token chunks, so the
[tokens, vocab] logits never exist in full.  The object returned as ``.loss`` is an autograd proxy on EVERY rank:
``loss.backward()`` runs that rank's part of the pipeline backward (SPMD equivalent of the autograd router).

Tied embeddings split over two ranks (Qwen2.5-0.5B at N > 1): the two copies' gradients are summed rank 0 <-> last
rank before the optimizer step, so both copies take identical updates.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .. import native as nat
from .module import CausalLMOutput

A_MN, B_MN, ACC = nat.A_MN_MAJOR, nat.B_MN_MAJOR, nat.EPI_ACCUM
HEAD_CHUNK = 2048          # tokens per fused lm_head + CE pass


class StageTrainer:
    """Training-mode execution of one ``CudaStage`` (activations saved per micro-batch)."""

    def __init__(self, stage):
        self.st = stage
        self.cfg = stage.cfg
        self.p = stage.params
        self.layer_ids = stage.params.layer_ids
        dev = stage.device
        names = [f"l{li}.{n}" for li in self.layer_ids for n in ("ln1", "ln2")] + (["norm"] if stage.has_head else [])
        self.norm_acc: Dict[str, torch.Tensor] = {n: torch.zeros(self.cfg.hidden, dtype=torch.float32, device=dev)
                                                  for n in names}
        if self.cfg.qkv_bias:
            for li in self.layer_ids:
                self.norm_acc[f"l{li}.bqkv"] = torch.zeros(self.cfg.qkv_dim, dtype=torch.float32, device=dev)
        if self.cfg.qk_norm:
            for li in self.layer_ids:
                for n in ("qn", "kn"):
                    self.norm_acc[f"l{li}.{n}"] = torch.zeros(self.cfg.head_dim, dtype=torch.float32, device=dev)
        self.grp = stage.slots[0]                     # rope tables / scale come from the shard operator
        self.ctx: Dict[int, dict] = {}
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        self.n_valid_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.launches = 0
        # Weight matrices whose gradient is produced by exactly one GEMM per micro-batch.  zero_grad() does not memset
        # them (99.9 % of the gradient arena): it marks them "fresh" and the first weight-gradient GEMM afterwards WRITES
        # the tensor instead of read-modify-writing zeros (saves the 15 GB memset and a 15 GB read per step at 7B).
        self._lazy = [f"l{li}.{n}" for li in self.layer_ids for n in ("wqkv", "wo", "wgu", "wd")]
        if stage.has_head and not (self.cfg.tied and stage.has_embed) and "head" in self.p.g and not self.cfg.tied:
            self._lazy.append("head")
        self._fresh: set = set()
        self._eager = [t for n, t in self.p.g.items() if n not in set(self._lazy)]
        self.p.grad_settle = self.settle_grads          # gradient export (hf_state_dict(grads=True)) settles first

    # ------------------------------------------------------------------------------------------ forward
    def forward_layers(self, mb: int, x: torch.Tensor) -> torch.Tensor:
        """x [b,S,H] -> [b,S,H] through this stage's layers, saving what the backward needs."""
        cfg, v = self.cfg, self.p.v
        b, S, H = x.shape
        N = b * S
        dev, bf = x.device, torch.bfloat16
        x = x.reshape(N, H).contiguous()
        saved: List[dict] = []
        zero_pos = torch.zeros(1, dtype=torch.int32, device=dev)
        for li in self.layer_ids:
            s = {"x_in": x}
            s["rstd1"] = torch.empty(N, dtype=torch.float32, device=dev)
            s["h1"] = nat.rmsnorm_fwd(x, v[f"l{li}.ln1"], cfg.rms_eps, rstd=s["rstd1"])
            qkv = nat.gemm(s["h1"], v[f"l{li}.wqkv"], bias=v.get(f"l{li}.bqkv"))
            s["q"] = torch.empty(N, cfg.q_dim, dtype=bf, device=dev)
            s["kc"] = torch.empty(b, cfg.n_kv_heads, S, cfg.head_dim, dtype=bf, device=dev)
            s["vc"] = torch.empty_like(s["kc"])
            if cfg.qk_norm:
                s["qkv"] = qkv                      # pre-norm q/k are needed by the q/k-norm backward
            nat.rope_kv_fwd(qkv, s["q"], s["kc"], s["vc"], zero_pos, self.grp.cos, self.grp.sin, v.get(f"l{li}.qn"),
                            v.get(f"l{li}.kn"), cfg.rms_eps, S, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim)
            s["attn"] = torch.empty(N, cfg.q_dim, dtype=bf, device=dev)
            s["lse"] = torch.empty(b, cfg.n_heads, S, dtype=torch.float32, device=dev)
            nat.attn_prefill_fwd(s["q"], s["kc"], s["vc"], s["attn"], s["lse"], b, S, 0, cfg.n_heads, cfg.n_kv_heads,
                                 cfg.head_dim, self.grp.scale)
            s["x_mid"] = nat.gemm(s["attn"], v[f"l{li}.wo"], residual=x)
            s["rstd2"] = torch.empty(N, dtype=torch.float32, device=dev)
            s["h2"] = nat.rmsnorm_fwd(s["x_mid"], v[f"l{li}.ln2"], cfg.rms_eps, rstd=s["rstd2"])
            s["gu"] = nat.gemm(s["h2"], v[f"l{li}.wgu"])
            s["act"] = torch.empty(N, cfg.intermediate, dtype=bf, device=dev)     # kept: the down-proj wgrad needs it
            nat.swiglu_fwd(s["gu"], s["act"])
            x = nat.gemm(s["act"], v[f"l{li}.wd"], residual=s["x_mid"])
            saved.append(s)
            self.launches += 10
        self.ctx[mb] = {"layers": saved, "b": b, "S": S}
        return x.view(b, S, H)

    def head_loss_and_grad(self, mb: int, x: torch.Tensor, shift_labels: torch.Tensor, inv_n: float) -> None:
        """Last stage: final norm + lm_head + shifted CE, fused with its own backward, chunked over tokens.
        Accumulates the loss sum, the lm_head / final-norm gradients, and stores d(loss)/d(x) for ``backward``."""
        cfg, v, g = self.cfg, self.p.v, self.p.g
        b, S, H = x.shape
        N = b * S
        x2 = x.reshape(N, H)
        labels = shift_labels.reshape(N).contiguous()
        dx = torch.empty_like(x2)
        for a in range(0, N, HEAD_CHUNK):
            e = min(N, a + HEAD_CHUNK)
            xc = x2[a:e]
            rstd = torch.empty(e - a, dtype=torch.float32, device=x.device)
            hn = nat.rmsnorm_fwd(xc, v["norm"], cfg.rms_eps, rstd=rstd)
            logits = nat.gemm(hn, v["head"])
            nat.ce_fwd_bwd(logits, labels[a:e], self.loss_sum, self.n_valid_dev, logits, inv_n)
            dhn = nat.gemm(logits, v["head"], flags=B_MN, N=H)                       # [n,V]·[V,H]
            nat.gemm(logits, hn, out=g["head"], flags=A_MN | B_MN | self._acc("head"), M=cfg.vocab, K=e - a, N=H)   # dW += dlogits^T·hn
            nat.rmsnorm_bwd(xc, v["norm"], dhn, rstd, dx[a:e], self.norm_acc["norm"])
            self.launches += 6
        self.ctx[mb]["dx_out"] = dx.view(b, S, H)

    # ------------------------------------------------------------------------------------------ backward
    def backward_layers(self, mb: int, dy: torch.Tensor) -> torch.Tensor:
        """dy [b,S,H] (gradient of this stage's output) -> gradient of its input; parameter grads accumulate."""
        cfg, v, g = self.cfg, self.p.v, self.p.g
        c = self.ctx.pop(mb)
        b, S = c["b"], c["S"]
        N, H = b * S, cfg.hidden
        dev, bf = dy.device, torch.bfloat16
        dy = dy.reshape(N, H).contiguous()
        ws = torch.empty(max(nat.attn_bwd_ws(b, S, cfg.n_heads), 16), dtype=torch.uint8, device=dev)
        for j in reversed(range(len(self.layer_ids))):
            li, s = self.layer_ids[j], c["layers"][j]
            # ---- MLP
            act = s["act"]
            d_act = nat.gemm(dy, v[f"l{li}.wd"], flags=B_MN, N=cfg.intermediate)         # dy·Wd
            nat.gemm(dy, act, out=g[f"l{li}.wd"], flags=A_MN | B_MN | self._acc(f"l{li}.wd"), M=H, K=N, N=cfg.intermediate)
            dgu = torch.empty_like(s["gu"])
            nat.swiglu_bwd(s["gu"], d_act, dgu)
            dh2 = nat.gemm(dgu, v[f"l{li}.wgu"], flags=B_MN, N=H)
            nat.gemm(dgu, s["h2"], out=g[f"l{li}.wgu"], flags=A_MN | B_MN | self._acc(f"l{li}.wgu"), M=2 * cfg.intermediate, K=N, N=H)
            d_xmid = torch.empty(N, H, dtype=bf, device=dev)
            nat.rmsnorm_bwd(s["x_mid"], v[f"l{li}.ln2"], dh2, s["rstd2"], d_xmid, self.norm_acc[f"l{li}.ln2"], dx_add=dy)
            # ---- attention
            d_attn = nat.gemm(d_xmid, v[f"l{li}.wo"], flags=B_MN, N=cfg.q_dim)
            nat.gemm(d_xmid, s["attn"], out=g[f"l{li}.wo"], flags=A_MN | B_MN | self._acc(f"l{li}.wo"), M=H, K=N, N=cfg.q_dim)
            dq = torch.empty(N, cfg.q_dim, dtype=bf, device=dev)
            dk = torch.empty(b, cfg.n_heads, S, cfg.head_dim, dtype=bf, device=dev)     # one partial per query head
            dv = torch.empty_like(dk)
            nat.attn_bwd(s["q"], s["kc"], s["vc"], s["attn"], d_attn, s["lse"], dq, dk, dv, ws, b, S, cfg.n_heads,
                         cfg.n_kv_heads, cfg.head_dim, self.grp.scale)
            dqkv = torch.empty(N, cfg.qkv_dim, dtype=bf, device=dev)
            nat.rope_kv_bwd(dq, dk, dv, dqkv, self.grp.cos, self.grp.sin, S, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim)
            if cfg.qk_norm:
                nat.qk_norm_bwd(s["qkv"], dqkv, v[f"l{li}.qn"], v[f"l{li}.kn"], self.norm_acc[f"l{li}.qn"],
                                self.norm_acc[f"l{li}.kn"], cfg.rms_eps, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim)
            dh1 = nat.gemm(dqkv, v[f"l{li}.wqkv"], flags=B_MN, N=H)
            nat.gemm(dqkv, s["h1"], out=g[f"l{li}.wqkv"], flags=A_MN | B_MN | self._acc(f"l{li}.wqkv"), M=cfg.qkv_dim, K=N, N=H)
            if cfg.qkv_bias:
                nat.colsum(dqkv, self.norm_acc[f"l{li}.bqkv"])
            dx = torch.empty(N, H, dtype=bf, device=dev)
            nat.rmsnorm_bwd(s["x_in"], v[f"l{li}.ln1"], dh1, s["rstd1"], dx, self.norm_acc[f"l{li}.ln1"], dx_add=d_xmid)
            dy = dx
            self.launches += 17
        return dy.view(b, S, H)

    def embed_backward(self, ids: torch.Tensor, dx: torch.Tensor):
        nat.embed_bwd(ids.reshape(-1).contiguous(), dx.reshape(-1, self.cfg.hidden).contiguous(), self.p.g["embed"])
        self.launches += 1

    def finish_backward(self):
        """fold the fp32 norm-gain accumulators into the bf16 gradient arena"""
        self.settle_grads()
        for n, acc in self.norm_acc.items():
            nat.f32_to_bf16_accum(acc, self.p.g[n], accumulate=True)
            acc.zero_()

    def _acc(self, name: str) -> int:
        """EPI_ACCUM unless this is the first gradient GEMM into ``name`` since zero_grad()."""
        if name in self._fresh:
            self._fresh.discard(name)
            return 0
        return ACC

    def settle_grads(self):
        """Give every still-fresh (never written since zero_grad) matrix gradient its zeros: called before anything reads
        the gradient arena as a whole (optimizer step, gradient export)."""
        for n in self._fresh:
            self.p.g[n].zero_()
        self._fresh.clear()

    def zero_grad(self):
        torch._foreach_zero_(self._eager)
        self._fresh = set(self._lazy)
        for a in self.norm_acc.values():
            a.zero_()


class _PipelineRouter(torch.autograd.Function):
    """SPMD counterpart of the reference's ``CustomAutogradRouter`` (module.py:126-144)."""

    @staticmethod
    def forward(ctx, anchor, dm, value):
        ctx.dm = dm
        return value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        train_backward(ctx.dm, float(grad_out))
        return torch.zeros(()), None, None


def _trainer(dm) -> StageTrainer:
    if getattr(dm.stage, "trainer", None) is None:
        make = getattr(dm.stage, "make_trainer", None)          # test backends provide their own twin
        dm.stage.trainer = make() if make else StageTrainer(dm.stage)
    return dm.stage.trainer


def train_forward(dm, input_ids: Optional[torch.Tensor], labels: Optional[torch.Tensor]) -> CausalLMOutput:
    """Forward of all micro-batches (GPipe order).  Every rank returns an autograd proxy as ``.loss``."""
    link, st, cfg, dev = dm.link, dm.stage, dm.cfg, dm.device
    tr = _trainer(dm)
    meta = None
    if link.first:
        if labels is None:
            raise ValueError("training forward needs labels= (the loss is produced on the last stage)")
        B, S = input_ids.shape
        shift = F.pad(labels, (0, 1), value=-100)[:, 1:].contiguous()
        meta = (B, S, int((shift != -100).sum()))
    B, S, n_valid = link.broadcast_object(meta)
    n_mb = min(dm.n_pipelines, B)
    if B % n_mb:
        raise ValueError(f"batch {B} not divisible into {n_mb} micro-batches")
    b = B // n_mb
    if link.first:
        ids_dev, shift_dev = input_ids.to(dev), shift.to(dev)
        if dm.world > 1:
            link.send_down(shift_dev, dm.world - 1)
    elif link.last:
        shift_dev = torch.empty(B, S, dtype=torch.int64, device=dev)
        link.recv_down(shift_dev, 0)
    tr.loss_sum.zero_()
    tr.n_valid_dev.zero_()
    dm._train_state = {"n_mb": n_mb, "b": b, "S": S, "ids": ids_dev if link.first else None}
    for m in range(n_mb):
        if link.first:
            x = st.embed(ids_dev[m * b:(m + 1) * b])
        else:
            x = torch.empty(b, S, cfg.hidden, dtype=torch.bfloat16, device=dev)
            link.recv_prev(x)
        x = tr.forward_layers(m, x)
        if not link.last:
            link.send_next(x.contiguous())
        else:
            tr.head_loss_and_grad(m, x, shift_dev[m * b:(m + 1) * b], 1.0 / max(n_valid, 1))
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    if link.last:
        loss = tr.loss_sum / max(n_valid, 1)
    if dm.world > 1:
        link.broadcast(loss, dm.world - 1)
    anchor = torch.zeros((), requires_grad=True)
    proxy = _PipelineRouter.apply(anchor, dm, loss.detach().float().cpu().reshape(()))
    return CausalLMOutput(loss=proxy, logits=None)


def train_backward(dm, grad_scale: float = 1.0):
    """Pipeline backward, micro-batches in reverse (module.py:414-524 / worker.py:233-295)."""
    link, st, cfg, dev = dm.link, dm.stage, dm.cfg, dm.device
    tr = _trainer(dm)
    s = dm._train_state
    for m in reversed(range(s["n_mb"])):
        if link.last:
            dy = tr.ctx[m].pop("dx_out")
            if grad_scale != 1.0:
                dy = dy * grad_scale
        else:
            dy = torch.empty(s["b"], s["S"], cfg.hidden, dtype=torch.bfloat16, device=dev)
            link.recv_next(dy)
        dx = tr.backward_layers(m, dy)
        if not link.first:
            link.send_prev(dx.contiguous())
        else:
            tr.embed_backward(s["ids"][m * s["b"]:(m + 1) * s["b"]], dx)
    link.flush()
    tr.finish_backward()
    if cfg.tied and dm.world > 1 and (link.first or link.last):
        # module.py:1218-1265 ties lm_head to embed_tokens on the host; here the two copies live on different ranks
        p = st.params
        if link.last:
            link.send_up(p.g["head"], 0)
            link.flush()
            link.recv_down(p.g["head"], 0)
        else:
            tmp = torch.empty_like(p.g["embed"])
            link.recv_up(tmp, dm.world - 1)
            nat.add_inplace(p.g["embed"].view(-1), tmp.view(-1))
            link.send_down(p.g["embed"], dm.world - 1)
            link.flush()


class StageAdam:
    """``create_optimizer(**kw)`` result: ``step()`` / ``zero_grad()`` over this rank's arena (optim.py:131-187)."""

    def __init__(self, dm, decoupled: bool = False, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, **_):
        self.dm, self.decoupled = dm, decoupled
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        p = dm.stage.params
        if p.grad is None:
            raise RuntimeError("DistributedModel was built with training=False; no gradient arena")
        self.m = torch.zeros(p.numel, dtype=torch.float32, device=p.device)
        self.v = torch.zeros_like(self.m)
        self.t = 0

    def zero_grad(self, set_to_none: bool = False):
        _trainer(self.dm).zero_grad()

    def step(self, closure=None):
        self.t += 1
        p = self.dm.stage.params
        nat.adamw_step(p.flat, p.grad, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t,
                       self.decoupled)
