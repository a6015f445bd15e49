"""TEST-ONLY stage backend: the CPU oracle behind the ``CudaStage`` interface.

Lets the multi-process host logic of ``DistributedModel`` (plan handling, micro-batch rotation, send/recv ordering,
label shipping, loss broadcast, backward routing) run under ``gloo`` on a CPU-only box.  It is injected through the
private ``_stage_factory`` hook from tests only; the product has no CPU path (``tensorlink_b200.native`` raises).
"""
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from oracle import shard_oracle as O
from tensorlink_b200.ml.weights import init_state_dict


class _Slot:
    def __init__(self, layer_ids):
        self.layer_ids = list(layer_ids)
        self.cache = O.KVCache()
        self.pos = 0


class OracleStage:
    def __init__(self, cfg, layer_ids, has_embed, has_head, device, max_batch, max_seq, n_slots=1, training=False,
                 state_dict=None, seed=1234, init="seeded", max_tokens=None):
        self.cfg, self.device = cfg, torch.device("cpu")
        self.has_embed, self.has_head = has_embed, has_head
        self.supports_training, self.trainer = bool(training), None
        self.layer_ids = list(layer_ids)
        sd = state_dict or init_state_dict(cfg, seed, torch.bfloat16, "cpu", layers=self.layer_ids,
                                           with_embed=has_embed or (has_head and cfg.tied), with_head=has_head)
        self.sd = {k: v.clone() for k, v in sd.items()}
        if training:
            for v in self.sd.values():
                v.requires_grad_(True)
            if cfg.tied and has_head and has_embed:
                self.sd["lm_head.weight"] = self.sd["model.embed_tokens.weight"]
        self.layers = {li: O.LayerWeights.from_state_dict(self.sd, li) for li in self.layer_ids}
        self.max_batch, self.max_seq = max_batch, max_seq
        self.slots: List[_Slot] = [_Slot(layer_ids) for _ in range(n_slots)]
        self.x_dec = [torch.zeros(max_batch, cfg.hidden, dtype=torch.bfloat16) for _ in range(n_slots)]
        self.ids_dec = [torch.zeros(max_batch, dtype=torch.int64) for _ in range(n_slots)]

    # ---- inference
    def embed(self, ids):
        return F.embedding(ids, self.sd["model.embed_tokens.weight"]).detach()

    def _run(self, x, slot, past):
        cfg = self.cfg
        B, S, _ = x.shape
        cos, sin = O.rope_tables(cfg, torch.arange(past, past + S)[None].expand(B, -1), x.dtype)
        with torch.no_grad():
            return O.shard_forward(cfg, [self.layers[i] for i in self.layer_ids], self.layer_ids, x, cos, sin,
                                   "sdpa_math", slot.cache)

    def prefill(self, hidden, past_len=0, slot=0):
        s = self.slots[slot]
        if past_len == 0:
            s.cache = O.KVCache()
        s.pos = past_len + hidden.shape[1]
        return self._run(hidden, s, past_len)

    def head_logits(self, hidden):
        with torch.no_grad():
            return F.linear(O.rmsnorm(hidden, self.sd["model.norm.weight"], self.cfg.rms_eps), self.sd["lm_head.weight"])

    def head_argmax(self, hidden, ids_out, slot=0):
        ids_out.copy_(self.head_logits(hidden).float().argmax(-1))

    def decode(self, slot, B, use_graph=True):
        s = self.slots[slot]
        x = self.x_dec[slot][:B]
        if self.has_embed:
            x.copy_(self.embed(self.ids_dec[slot][:B]))
        y = self._run(x[:, None, :], s, s.pos)[:, 0]
        s.pos += 1
        x.copy_(y)
        if self.has_head:
            self.head_argmax(x, self.ids_dec[slot][:B])

    def n_decode_launches(self, B):
        return 0

    def make_trainer(self):
        return OracleTrainer(self)

    @property
    def params(self):
        return self

    def hf_state_dict(self, grads: bool = False):
        return {k: (v.grad if grads else v.detach()) for k, v in self.sd.items() if not (grads and v.grad is None)}

    @property
    def g(self):
        """gradient views under the arena names the tied-embedding exchange uses (ml/train.py:train_backward)"""
        out = {}
        if "model.embed_tokens.weight" in self.sd:
            out["embed"] = self.sd["model.embed_tokens.weight"].grad
        if "lm_head.weight" in self.sd:
            out["head"] = self.sd["lm_head.weight"].grad
        return out


class OracleTrainer:
    """torch-autograd twin of ``StageTrainer`` (same method contract)."""

    def __init__(self, st: OracleStage):
        self.st, self.cfg = st, st.cfg
        self.ctx: Dict[int, dict] = {}
        self.loss_sum = torch.zeros(1)
        self.n_valid_dev = torch.zeros(1, dtype=torch.int32)
        self.launches = 0

    def forward_layers(self, mb, x):
        cfg, st = self.cfg, self.st
        b, S, _ = x.shape
        xin = x.detach().clone().requires_grad_(True)
        cos, sin = O.rope_tables(cfg, torch.arange(S)[None].expand(b, -1), x.dtype)
        y = O.shard_forward(cfg, [st.layers[i] for i in st.layer_ids], st.layer_ids, xin, cos, sin, "sdpa_math")
        self.ctx[mb] = {"xin": xin, "y": y, "b": b, "S": S}
        return y.detach()

    def head_loss_and_grad(self, mb, x, shift_labels, inv_n):
        st, cfg = self.st, self.cfg
        c = self.ctx[mb]
        y = c["y"]                       # keep the graph: head loss backward flows into the layers later
        logits = F.linear(O.rmsnorm(y, st.sd["model.norm.weight"], cfg.rms_eps), st.sd["lm_head.weight"]).float()
        ls = F.cross_entropy(logits.reshape(-1, cfg.vocab), shift_labels.reshape(-1), ignore_index=-100, reduction="sum")
        self.loss_sum += ls.detach()
        c["head_loss"] = ls * inv_n
        c["dx_out"] = torch.zeros_like(x)   # placeholder: the real gradient flows through the retained graph

    def backward_layers(self, mb, dy):
        c = self.ctx.pop(mb)
        if "head_loss" in c:
            c["head_loss"].backward()
        else:
            c["y"].backward(dy)
        g = c["xin"].grad
        return g if g is not None else torch.zeros_like(c["xin"])      # worker.py:274-277: missing grad -> zeros

    def embed_backward(self, ids, dx):
        e = self.st.sd["model.embed_tokens.weight"]
        with torch.enable_grad():          # we are inside an autograd backward (grad mode off)
            F.embedding(ids, e).backward(dx)

    def finish_backward(self):
        pass

    def zero_grad(self):
        for v in self.st.sd.values():
            v.grad = None
