"""``DistributedModel`` — the reference's user-facing front end, driving B200 pipeline stages.

Mirrors /root/reference/tensorlink/ml/module.py: constructor signature (:251-265), ``forward`` returning an HF-style
output with ``.logits`` / ``.loss`` (:348-407), ``generate`` (:763-769), ``create_optimizer`` (:1016-1021),
``train/eval`` (:534-566), ``parameters`` (:577-650), ``distribute_model(config)`` (:699).  The reference is
hub-and-spoke: the user process RPCs each worker in turn over TCP and polls.  Here every pipeline stage is one
process on one GPU (``torchrun``), all of them construct the same ``DistributedModel`` (SPMD) and activations
go stage -> stage directly over NVLink (p2p/link.py).  With one process the whole model is a single stage.

Documented deviations from the reference: errors raise instead of being swallowed into ``{"error": ...}`` dicts
(module.py:978-985); ``dtype`` defaults to bf16 (the only dtype the sm_100a kernels implement); logits live on
the last stage (pass ``gather_logits=True`` to copy them to rank 0).
"""
from __future__ import annotations

import os

import time
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Union

import torch

from ..p2p.link import StageLink, init_process_group_from_env
from . import graphing
from .configs import ShardModelConfig, get_config


@dataclass
class CausalLMOutput:
    """The two fields of HF ``CausalLMOutputWithPast`` the reference's callers read."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        return getattr(self, k)


def _default_stage_factory(**kw):
    from .stage import CudaStage          # imports the CUDA library; raises without it
    return CudaStage(**kw)


def _config_from_hf(model) -> ShardModelConfig:
    c = model.config
    qk_norm = c.__class__.__name__.startswith("Qwen3")
    hd = getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads
    rp = getattr(c, "rope_parameters", None) or {}
    theta = rp.get("rope_theta", getattr(c, "rope_theta", 1e6))
    return ShardModelConfig(getattr(c, "name_or_path", "") or c.__class__.__name__, c.hidden_size,
                            c.intermediate_size, c.num_hidden_layers, c.num_attention_heads,
                            c.num_key_value_heads, hd, c.vocab_size, tied=bool(c.tie_word_embeddings),
                            qkv_bias=not qk_norm, qk_norm=qk_norm, rope_theta=float(theta),
                            rms_eps=float(c.rms_norm_eps), max_pos=int(c.max_position_embeddings))


# HF forward / generate keywords that change nothing here when left at these values (anything else raises instead of
# being dropped silently: the reference forwards every keyword to HF, module.py:763-769)
_NEUTRAL_KW = {"use_cache": (True, False, None), "return_dict": (True, None), "output_attentions": (False, None),
               "output_hidden_states": (False, None), "num_beams": (1, None), "num_return_sequences": (1, None),
               "repetition_penalty": (1.0, None), "past_key_values": (None,), "position_ids": (None,),
               "return_dict_in_generate": (False, None), "logits_to_keep": (0, None), "min_new_tokens": (0, None)}


def _check_unconsumed(kwargs: dict, what: str):
    for k, v in kwargs.items():
        ok = _NEUTRAL_KW.get(k)
        if ok is None or not any(v is o or (o is not None and v == o) for o in ok):
            raise NotImplementedError(f"{what}: keyword {k}={v!r} is not supported by the B200 stage executor "
                                      "(it would be silently ignored otherwise)")


def _check_attention_mask(mask, shape):
    """forward(): only the all-ones mask (no padding) is accepted; a mask with zeros means padded prompts, whose rows
    would otherwise attend to pad tokens at shifted positions.  (``generate`` handles left-padded batches.)"""
    if mask is None:
        return
    if tuple(mask.shape) != tuple(shape):
        raise ValueError(f"attention_mask shape {tuple(mask.shape)} != input_ids shape {tuple(shape)}")
    if not bool((mask != 0).all()):
        raise NotImplementedError("padded rows (attention_mask with zeros) are not supported by forward(): pass rows of "
                                  "equal length")


def _left_pad_groups(mask: torch.Tensor):
    """A left-padded batch (HF's convention for generation: ``tokenizer(..., padding=True, padding_side='left')``)
    as {real length: [row indices]}; None when no row is padded.  Right padding / holes raise."""
    m = (mask != 0).cpu()
    if bool(m.all()):
        return None
    S = m.shape[1]
    lengths = m.sum(1)
    want = torch.arange(S)[None, :] >= (S - lengths)[:, None]           # zeros, then ones
    if not torch.equal(m, want) or int(lengths.min()) == 0:
        raise NotImplementedError("attention_mask must describe LEFT-padded prompts (zeros, then ones; at least one token per row)")
    groups = {}
    for r, L in enumerate(lengths.tolist()):
        groups.setdefault(int(L), []).append(r)
    return groups


EOS_CHECK_EVERY = 16       # decode steps between two host-side "has every row emitted EOS?" checks


def _eos_list(eos_token_id):
    if eos_token_id is None:
        return []
    if isinstance(eos_token_id, torch.Tensor):              # HF accepts an int, a list or a tensor of ids
        return [int(e) for e in eos_token_id.reshape(-1).tolist()]
    return [int(e) for e in eos_token_id] if isinstance(eos_token_id, (list, tuple)) else [int(eos_token_id)]


def _all_rows_finished(tokens: torch.Tensor, eos_ids) -> bool:
    """tokens [rows, cols] generated so far: True once every row holds an EOS (one small device reduction + sync)."""
    hit = torch.zeros_like(tokens, dtype=torch.bool)
    for e in eos_ids:
        hit |= tokens == e
    return bool(hit.any(1).all().item())


def apply_eos(result: torch.Tensor, prompt_len: int, eos_token_id=None, pad_token_id=None) -> torch.Tensor:
    """HF ``generate`` stopping semantics applied to a generation [B, S+new]: everything after a row's first EOS
    becomes ``pad_token_id`` (default: the EOS id) and the result ends where the last row finished.  (The decode loop
    checks every ``EOS_CHECK_EVERY`` steps whether all rows are finished and stops computing then; the at most 15 surplus
    columns are cut here.)"""
    if eos_token_id is None:
        return result
    eos_ids = _eos_list(eos_token_id)
    pad = eos_ids[0] if pad_token_id is None else int(pad_token_id)
    new = result[:, prompt_len:]
    if new.shape[1] == 0:
        return result
    is_eos = torch.zeros_like(new, dtype=torch.bool)
    for e in eos_ids:
        is_eos |= new == e
    seen_before = (is_eos.cumsum(1) - is_eos.long()) > 0
    new = torch.where(seen_before, torch.full_like(new, pad), new)
    first = torch.where(is_eos.any(1), is_eos.long().argmax(1) + 1, torch.full_like(new[:, 0], new.shape[1]))
    return torch.cat([result[:, :prompt_len], new[:, :int(first.max())]], dim=1)


class DistributedModel(torch.nn.Module):
    def __init__(self, model: Union[torch.nn.Module, str, ShardModelConfig], n_pipelines: int = 1,
                 optimizer=None, scheduler_type=None, device: Optional[str] = None,
                 dtype: torch.dtype = torch.bfloat16, trusted: bool = False, node: Optional[Any] = None,
                 training: bool = True, verbose: bool = False, tokenizer=None, config: Optional[dict] = None,
                 *, max_batch: int = 8, max_seq: int = 4096, seed: int = 1234, init: str = "seeded",
                 balanced_plan: bool = False, link: Optional[StageLink] = None, max_tokens: Optional[int] = None,
                 _stage_factory: Optional[Callable] = None):
        super().__init__()
        if dtype != torch.bfloat16:
            raise NotImplementedError("tensorlink_b200 computes in bf16 with fp32 accumulation; pass dtype=torch.bfloat16")
        state_dict = None
        if isinstance(model, torch.nn.Module):
            self.cfg = _config_from_hf(model)
            state_dict = model.state_dict()
        elif isinstance(model, ShardModelConfig):
            self.cfg = model
        elif isinstance(model, str) and os.path.isdir(model) and os.path.exists(os.path.join(model, "config.json")):
            # a local checkpoint in the HF layout: each stage reads only its own tensors (worker.py:542-638)
            from .checkpoint import LazyCheckpoint, config_from_dir
            self.cfg = config_from_dir(model)
            state_dict = LazyCheckpoint(model)
        else:
            self.cfg = get_config(model)
        self.model_name = self.cfg.name
        self.name = self.model_name
        self.tokenizer = tokenizer
        self.n_pipelines = max(1, int(n_pipelines))          # micro-batches in flight (module.py:374-399)
        self.n_datalines = 1
        self.optimizer = optimizer
        self.scheduler = scheduler_type
        self.training = training
        self.verbose = verbose
        self.trusted = trusted
        self.job_id = None
        if node is None:
            from ..nodes.nodes import User
            node = User()
        self.node = node
        self.node_requests = getattr(node, "node_requests", None)
        self.node_responses = getattr(node, "node_responses", None)
        self.mpc_lock = getattr(node, "mpc_lock", None)

        if link is None:
            init_process_group_from_env()
            link = StageLink.from_env()
        self.link = link
        self.rank, self.world = link.rank, link.world
        if device is None:
            device = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)

        self.config = config if config else {}
        self.max_batch, self.max_seq, self.seed = max_batch, max_seq, seed
        self._stage_factory = _stage_factory or _default_stage_factory
        self._stage_kw = dict(init=init, state_dict=state_dict, balanced=balanced_plan, max_tokens=max_tokens)
        self.distributed_graph: Dict[str, Any] = {}
        self.stage = None
        self.timers: Dict[str, float] = {}
        if self.node.__class__.__name__ == "User":
            self._initialize_distribution()

    # ------------------------------------------------------------------------------------------ distribution
    def _initialize_distribution(self):
        """module.py:987-1021: obtain a plan, distribute, expose ``create_optimizer``."""
        plan = self.config or graphing.make_plan(self.cfg, self.world, self.training, self._stage_kw["balanced"])
        self.distribute_model(plan)

    def distribute_model(self, config: Optional[dict] = None):
        """module.py:699-761.  Every rank materialises exactly its entries of the plan."""
        plan = config or self.config or graphing.make_plan(self.cfg, self.world, self.training)
        if graphing.n_stages(plan) != self.world:
            raise ValueError(f"plan has {graphing.n_stages(plan)} stages but the job has {self.world} ranks")
        self.distributed_graph = plan
        layers = graphing.stage_layers(plan, self.rank)
        n_slots = self.n_pipelines
        per_slot = (self.max_batch + n_slots - 1) // n_slots
        self.stage = self._stage_factory(cfg=self.cfg, layer_ids=layers, has_embed=self.rank == 0,
                                         has_head=self.rank == self.world - 1, device=self.device,
                                         max_batch=per_slot, max_seq=self.max_seq, n_slots=n_slots,
                                         training=self.training, state_dict=self._stage_kw["state_dict"],
                                         seed=self.seed, init=self._stage_kw["init"],
                                         max_tokens=self._stage_kw["max_tokens"])
        self._stage_kw["state_dict"] = None
        return plan

    # ------------------------------------------------------------------------------------------ nn.Module surface
    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def parameters(self, recurse: bool = True, distributed: bool = True, load: bool = True):
        """module.py:577-650: this rank's parameters under their HF names (values, not nn.Parameters)."""
        return iter(self.stage.params.hf_state_dict().values())

    def state_dict(self, *a, gather: bool = False, **k):
        """This rank's tensors under their HF names.  ``gather=True``: the WHOLE model's state dict on the first rank, on the
        host (what the reference's ``parameters(distributed=True, load=True)`` pulls from its workers, module.py:577-650);
        the other ranks get their own part.  For checkpoints prefer ``save_pretrained`` (each stage writes its own file)."""
        sd = self.stage.params.hf_state_dict()
        if not gather or self.world == 1:
            return sd
        parts = self.link.gather_object({k_: v.detach().cpu() for k_, v in sd.items()}, 0)
        if not self.link.first:
            return sd
        whole: Dict[str, torch.Tensor] = {}
        for part in parts:
            whole.update(part)
        if self.cfg.tied and "lm_head.weight" in whole and "model.embed_tokens.weight" in whole:
            whole["lm_head.weight"] = whole["model.embed_tokens.weight"]      # one tensor under both names, like HF
        return whole

    def save_pretrained(self, path: str):
        """Write this job's weights as an HF-layout checkpoint (one safetensors file per stage + index + config)."""
        from .checkpoint import save_checkpoint
        save_checkpoint(self, path)

    def create_optimizer(self, **optimizer_kwargs):
        from .optim import create_distributed_optimizer
        return create_distributed_optimizer(self, self.optimizer, **optimizer_kwargs)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, *args, **kwargs) -> CausalLMOutput:
        """One forward through every stage (module.py:348-407).  ``input_ids`` positional or keyword (:355-359).
        Inference: logits [B,S,V] on the last stage.  Training (``self.training`` with a grad-enabled stage):
        handled by ``ml/train.py``."""
        input_ids = kwargs.pop("input_ids", args[0] if args else None)
        labels = kwargs.pop("labels", None)
        gather = kwargs.pop("gather_logits", False)
        if self.link.first and input_ids is not None:
            _check_attention_mask(kwargs.pop("attention_mask", None), input_ids.shape)
        else:
            kwargs.pop("attention_mask", None)
        _check_unconsumed(kwargs, "DistributedModel.forward")
        if self.training and getattr(self.stage, "supports_training", False):
            from .train import train_forward
            return train_forward(self, input_ids, labels)
        return self._infer_forward(input_ids, gather)

    def _infer_forward(self, input_ids: Optional[torch.Tensor], gather: bool) -> CausalLMOutput:
        link, st, cfg = self.link, self.stage, self.cfg
        shape = link.broadcast_object(tuple(input_ids.shape) if link.first else None)
        B, S = shape
        # micro-batches of at most the stage's per-slot batch flow through the ranks back to back
        mb = min(B, st.max_batch)
        parts = []
        for a in range(0, B, mb):
            e = min(B, a + mb)
            if link.first:
                x = st.embed(input_ids[a:e].to(self.device))
            else:
                x = torch.empty(e - a, S, cfg.hidden, dtype=torch.bfloat16, device=self.device)
                link.recv_prev(x)
            x = st.prefill(x, 0, 0)
            if not link.last:
                link.send_next(x.clone())
            else:
                parts.append(st.head_logits(x.reshape((e - a) * S, cfg.hidden)).view(e - a, S, cfg.vocab))
        logits = torch.cat(parts, dim=0) if parts else None
        if gather and self.world > 1:
            if link.last:
                link.send_up(logits.contiguous(), 0)
            elif link.first:
                logits = torch.empty(B, S, cfg.vocab, dtype=torch.bfloat16, device=self.device)
                link.recv_up(logits, self.world - 1)
        link.flush()
        return CausalLMOutput(logits=logits)

    # ------------------------------------------------------------------------------------------ peer-memory decode
    def _peer_ring(self, n_mb: int):
        """The mailbox ring for decode hops (p2p/peer.py), or None: TL_P2P=nccl, or a stage that is not the CUDA one
        (the gloo tests drive this class with a CPU stage).  Built collectively on first use."""
        import os
        from .stage import CudaStage
        if os.environ.get("TL_P2P", "peer") == "nccl" or not isinstance(self.stage, CudaStage):
            return None
        if getattr(self, "_ring", None) is None:
            from ..p2p.peer import PeerRing
            st = self.stage
            self._ring = PeerRing(self.link, len(st.slots), st.max_batch, self.cfg.hidden, st.max_seq, self.device)
        return self._ring

    def _decode_ring(self, ring, input_ids, B, S, b, n_mb, max_new, streamer, use_graph, profile, t0):
        """Decode rounds with every hop on peer memory.  The host only enqueues: max_new-1 graph replays per
        micro-batch (wait -> layers -> store into the neighbour -> signal), no synchronisation until the end; the
        first stage logs each token column on the device (``ring.out_log``)."""
        link, st, dev = self.link, self.stage, self.device
        span = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        span[0].record()
        step_done = []
        eos_ids = _eos_list(self._eos[0])
        n_cols = max_new                              # token columns that will exist when the loop ends
        for step in range(max_new):
            for m in range(n_mb):
                if step < max_new - 1:
                    st.decode(m, b, use_graph, ring=ring)
                elif link.first:                                   # last column: nothing left to compute
                    ring.wait_ids(m)
                    ring.log_token(m, b)
            if streamer is not None and link.first:
                ev = torch.cuda.Event()
                ev.record()
                step_done.append(ev)
            if eos_ids and step < max_new - 1 and (step + 1) % EOS_CHECK_EVERY == 0:
                # HF stops once every row has emitted EOS.  Columns 0..step are logged on the first stage; all ranks
                # drain their queues (no persistent kernel is in flight during the collective) and agree on stopping.
                torch.cuda.synchronize(dev)
                flag = torch.zeros(1, dtype=torch.int32, device=dev)
                if link.first and _all_rows_finished(ring.out_log[:n_mb, :b, :step + 1].reshape(B, step + 1), eos_ids):
                    flag.fill_(1)
                link.broadcast(flag, 0)
                if int(flag.item()):
                    n_cols = step + 1
                    break
        span[1].record()
        if streamer is not None and link.first:
            for step, ev in enumerate(step_done):                  # step's graph logged column `step` before its event
                ev.synchronize()
                streamer.put(ring.out_log[:n_mb, :b, step].reshape(-1).cpu())
        torch.cuda.synchronize(dev)
        ring.check()
        if hasattr(st, "check"):
            st.check()
        if profile:
            self.timers["decode_span_s"] = span[0].elapsed_time(span[1]) * 1e-3
            self.timers["decode_busy_s"] = self.timers["decode_span_s"] - float(ring.wait_ns.item()) * 1e-9
        if link.first:
            out_tokens = ring.out_log[:n_mb, :b, :n_cols].reshape(B, n_cols)
            result = torch.cat([input_ids.to(dev), out_tokens], dim=1)
        else:
            result = torch.empty(B, S + n_cols, dtype=torch.int64, device=dev)
        link.broadcast(result, 0)
        if streamer is not None and link.first:
            streamer.end()
        self.timers["generate_wall_s"] = time.perf_counter() - t0
        return apply_eos(result, S, *self._eos)

    def _generate_left_padded(self, input_ids, groups, shape, max_new, streamer, use_graph, sampling):
        """HF semantics for a left-padded batch: every row attends to its own tokens only, at positions 0..L-1.  Rows of
        equal real length are generated together (one uniform run per length: the KV cache and RoPE positions of a run
        start at the row's first real token, so no pad key exists to be masked); the result keeps HF's layout
        [pads | prompt | new tokens | pad_token_id...]."""
        if streamer is not None:
            raise NotImplementedError("streamer with a padded batch (rows finish in separate runs)")
        eos, pad = self._eos
        pad_id = pad if pad is not None else (_eos_list(eos)[0] if eos is not None else 0)
        first = self.link.first
        B, S = shape
        out = torch.full((B, S + max_new), int(pad_id), dtype=torch.int64, device=self.device)
        if first:
            out[:, :S] = input_ids.to(self.device)
        longest = 0
        for L in sorted(groups):
            rows = groups[L]
            sub = input_ids[rows][:, S - L:].contiguous() if first else None
            kw = dict(max_new_tokens=max_new, use_graph=use_graph, eos_token_id=eos, pad_token_id=pad)
            if sampling is not None:
                kw.update(do_sample=True, temperature=sampling["temperature"], top_k=sampling["top_k"], top_p=sampling["top_p"],
                          seed=sampling["seed"] + L)
            got = self.generate(sub, **kw)
            n_new = got.shape[1] - L
            out[torch.as_tensor(rows, device=self.device), S:S + n_new] = got[:, L:].to(self.device)
            longest = max(longest, n_new)
        self._eos = (eos, pad)
        out = out[:, :S + longest].contiguous()
        if self.world > 1:
            self.link.broadcast(out, 0)              # every rank returns the whole result, prompt and pads included
        return out

    # ------------------------------------------------------------------------------------------ generate
    @torch.no_grad()
    def generate(self, *args, **kwargs) -> Optional[torch.Tensor]:
        """Generation (module.py:763-769 delegates to HF ``generate``).  Greedy by default; ``do_sample=True`` draws every
        token on the last stage's GPU from the distribution HF's warpers define — ``temperature`` (default 1.0), ``top_k``
        (default 50, 0 = off), ``top_p`` (default 1.0) — with a counter-based Philox stream keyed by ``seed`` (extension;
        default ``torch.initial_seed()``): the same seed reproduces the same tokens (csrc/sample.cu).
        ``input_ids`` [B,S] int64 on the first stage; returns [B,S+new] on every rank.
        ``streamer``: object with ``put(tensor)`` / ``end()`` (HF BaseStreamer protocol), called on rank 0
        with each new token column, all batch rows (the reference streams row 0 only, worker.py:134-139)."""
        input_ids = kwargs.pop("input_ids", args[0] if args else None)
        max_new = int(kwargs.pop("max_new_tokens", 20))
        streamer = kwargs.pop("streamer", None)
        use_graph = kwargs.pop("use_graph", True)
        profile = kwargs.pop("profile", False)       # CUDA events around every decode launch -> self.timers["decode_busy_s"]
        sampling = None
        temperature, top_k, top_p = kwargs.pop("temperature", None), kwargs.pop("top_k", None), kwargs.pop("top_p", None)
        seed = kwargs.pop("seed", None)
        if kwargs.pop("do_sample", False):
            sampling = {"temperature": 1.0 if temperature is None else float(temperature), "top_k": 50 if top_k is None else int(top_k),
                        "top_p": 1.0 if top_p is None else float(top_p), "seed": int(torch.initial_seed() if seed is None else seed)}
            if sampling["temperature"] <= 0 or not (0 < sampling["top_p"] <= 1) or sampling["top_k"] < 0:
                raise ValueError(f"invalid sampling parameters {sampling}")
        self._eos = (kwargs.pop("eos_token_id", None), kwargs.pop("pad_token_id", None))
        mask = kwargs.pop("attention_mask", None)
        _check_unconsumed(kwargs, "DistributedModel.generate")
        link, st, cfg = self.link, self.stage, self.cfg
        groups = None
        if link.first and mask is not None:
            if tuple(mask.shape) != tuple(input_ids.shape):
                raise ValueError(f"attention_mask shape {tuple(mask.shape)} != input_ids shape {tuple(input_ids.shape)}")
            g = _left_pad_groups(mask)
            groups = None if g is None else (g, tuple(input_ids.shape))
        if self.world > 1:
            groups = link.broadcast_object(groups)
        if groups is not None:
            return self._generate_left_padded(input_ids, groups[0], groups[1], max_new, streamer, use_graph, sampling)
        shape, sampling = link.broadcast_object((tuple(input_ids.shape), sampling) if link.first else None)
        B, S = shape
        if hasattr(st, "set_sampling"):
            st.set_sampling(sampling)               # the last stage draws; greedy (None) restores the argmax path
            if sampling is not None and st.has_head:
                st.sample_ctr.zero_()               # a seed names ONE stream: the same call reproduces its tokens
        elif sampling is not None:
            raise NotImplementedError("sampling needs the CUDA stage")
        n_mb = min(self.n_pipelines, B)
        while B % n_mb:                      # the largest micro-batch count <= n_pipelines that divides the batch
            n_mb -= 1
        b = B // n_mb
        if b > st.max_batch or S + max_new > st.max_seq:
            raise ValueError(f"stage sized for micro-batch<={st.max_batch}, T<={st.max_seq}; got {b}, {S + max_new}")
        dev = self.device
        t0 = time.perf_counter()
        out_tokens = torch.zeros(B, max_new, dtype=torch.int64, device=dev) if link.first else None
        ids_rows = [input_ids[m * b:(m + 1) * b].to(dev) for m in range(n_mb)] if link.first else [None] * n_mb

        # ---- prefill every micro-batch through the pipeline; the last stage produces the first new token
        multi = self.world > 1
        ring = self._peer_ring(n_mb) if multi else None
        if multi and ring is None and hasattr(st, "slots"):
            for g in st.slots:                       # NCCL kernels share the SMs during decode: no persistent all-SM kernel
                if hasattr(g, "allow_chain"):
                    g.allow_chain = False
        if ring is not None:
            if max_new > ring.max_new:
                raise ValueError(f"max_new_tokens {max_new} exceeds the token log of the peer ring ({ring.max_new})")
            ring.reset()
        for m in range(n_mb):
            if link.first:
                x = st.embed(ids_rows[m])
            else:
                x = torch.empty(b, S, cfg.hidden, dtype=torch.bfloat16, device=dev)
                link.recv_prev(x)
            x = st.prefill(x, 0, m)
            if not link.last:
                link.send_next(x.clone())
            elif ring is not None:                         # first token straight into the first stage's mailbox
                st.head_argmax(x[:, -1, :].contiguous(), ring.first_ids_in[m][:b], m)
                ring.signal_ids(m)
            else:
                st.head_argmax(x[:, -1, :].contiguous(), st.ids_dec[m][:b], m)
                if multi:
                    link.send_up(st.ids_dec[m][:b].clone(), 0)
        if ring is not None:
            return self._decode_ring(ring, input_ids, B, S, b, n_mb, max_new, streamer, use_graph, profile, t0)
        # ---- decode rounds: micro-batches rotate through the stages; hidden [b,H] hops down, ids hop back up.
        # Sends are asynchronous; a slot's buffer is waited on only right before the next step overwrites it.
        sent_x = [None] * n_mb
        sent_ids = [None] * n_mb
        prof_events = []
        if profile:
            span = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            span[0].record()
        eos_ids = _eos_list(self._eos[0])
        n_cols = max_new

        def collect(m, step):
            """first stage: ids of micro-batch m for column ``step`` (sent by the last stage in the previous round)"""
            if multi:
                link.wait(sent_x[m])                       # x_dec[m] still feeding the previous hop?
                link.recv_up(st.ids_dec[m][:b], self.world - 1)
            out_tokens[m * b:(m + 1) * b, step] = st.ids_dec[m][:b]

        for step in range(max_new):
            collected = False
            if eos_ids and step and step % EOS_CHECK_EVERY == 0:
                # Stop once every row has emitted EOS (HF semantics).  The first stage takes this column's ids of EVERY
                # micro-batch first, so that no send is left without its posted receive when the ranks meet in the
                # broadcast below (a collective queued behind an unmatched point-to-point op could wait forever).
                if link.first:
                    for m in range(n_mb):
                        collect(m, step)
                collected = True
                flag = torch.zeros(1, dtype=torch.int32, device=dev)
                if link.first and _all_rows_finished(out_tokens[:, :step + 1], eos_ids):
                    flag.fill_(1)
                if multi:
                    link.flush()
                    if torch.device(dev).type == "cuda":
                        torch.cuda.synchronize(dev)
                    link.broadcast(flag, 0)
                if int(flag.item()):
                    n_cols = step + 1
                    if streamer is not None and link.first:
                        streamer.put(out_tokens[:, step].cpu())
                    break
            for m in range(n_mb):
                if link.first:
                    if not collected:
                        collect(m, step)
                    if streamer is not None and m == n_mb - 1:
                        streamer.put(out_tokens[:, step].cpu())         # all rows of this step, one column
                if step == max_new - 1:
                    continue
                if not link.first:
                    link.wait(sent_x[m])
                    link.recv_prev(st.x_dec[m][:b])
                if link.last and multi:
                    link.wait(sent_ids[m])
                if profile:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                st.decode(m, b, use_graph)
                if profile:
                    ev[1].record()
                    prof_events.append(ev)
                if not link.last:
                    sent_x[m] = link.send_next(st.x_dec[m][:b])
                elif multi:
                    sent_ids[m] = link.send_up(st.ids_dec[m][:b], 0)
        link.flush()
        if profile:
            span[1].record()
            torch.cuda.synchronize()
            self.timers["decode_span_s"] = span[0].elapsed_time(span[1]) * 1e-3
            self.timers["decode_busy_s"] = sum(a.elapsed_time(b_) for a, b_ in prof_events) * 1e-3
        if link.first:
            result = torch.cat([input_ids.to(dev), out_tokens[:, :n_cols]], dim=1)
        else:
            result = torch.empty(B, S + n_cols, dtype=torch.int64, device=dev)
        link.broadcast(result, 0)
        if hasattr(st, "check"):
            st.check()
        if streamer is not None and link.first:
            streamer.end()
        self.timers["generate_wall_s"] = time.perf_counter() - t0
        return apply_eos(result, S, *self._eos)
