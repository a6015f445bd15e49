"""Checkpoint in / out in the Hugging Face layout.

The reference reads, per worker, only the tensors of its layer range out of the hub snapshot's safetensors files and
remaps their keys to the local module (/root/reference/tensorlink/ml/worker.py:542-638, key remap :602-606), and
gathers parameters back through ``parameters(distributed=True)`` (ml/module.py:577-650).  Here a stage asks a
``LazyCheckpoint`` for exactly the HF tensor names it owns — nothing else is read from disk — and
``save_checkpoint`` writes one ``model-XXXXX-of-YYYYY.safetensors`` per stage plus the index and ``config.json``, so
the directory loads back here, in the reference, or in ``transformers``.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterator

import torch

from .configs import ShardModelConfig

INDEX = "model.safetensors.index.json"
SINGLE = "model.safetensors"


def config_from_dir(path: str) -> ShardModelConfig:
    """``config.json`` (HF Qwen2 / Qwen3 causal LM) -> ShardModelConfig."""
    with open(os.path.join(path, "config.json")) as f:
        c = json.load(f)
    mt = c.get("model_type", "")
    if mt not in ("qwen2", "qwen3"):
        raise ValueError(f"{path}: model_type {mt!r} is not supported (qwen2 / qwen3)")
    qk_norm = mt == "qwen3"
    n_h = int(c["num_attention_heads"])
    hd = int(c.get("head_dim") or c["hidden_size"] // n_h)
    theta = (c.get("rope_parameters") or {}).get("rope_theta", c.get("rope_theta", 1e6))
    return ShardModelConfig(c.get("_name_or_path") or os.path.basename(os.path.normpath(path)), int(c["hidden_size"]),
                            int(c["intermediate_size"]), int(c["num_hidden_layers"]), n_h, int(c["num_key_value_heads"]), hd,
                            int(c["vocab_size"]), tied=bool(c.get("tie_word_embeddings", False)), qkv_bias=not qk_norm,
                            qk_norm=qk_norm, rope_theta=float(theta), rms_eps=float(c.get("rms_norm_eps", 1e-6)),
                            max_pos=int(c.get("max_position_embeddings", 32768)))


def config_to_json(cfg: ShardModelConfig) -> dict:
    return {"architectures": ["Qwen3ForCausalLM" if cfg.qk_norm else "Qwen2ForCausalLM"],
            "model_type": "qwen3" if cfg.qk_norm else "qwen2", "hidden_size": cfg.hidden,
            "intermediate_size": cfg.intermediate, "num_hidden_layers": cfg.n_layers, "num_attention_heads": cfg.n_heads,
            "num_key_value_heads": cfg.n_kv_heads, "head_dim": cfg.head_dim, "vocab_size": cfg.vocab,
            "tie_word_embeddings": bool(cfg.tied), "rope_theta": cfg.rope_theta, "rms_norm_eps": cfg.rms_eps,
            "max_position_embeddings": cfg.max_pos, "hidden_act": "silu", "torch_dtype": "bfloat16",
            "attention_bias": bool(cfg.qkv_bias), "_name_or_path": cfg.name}


class LazyCheckpoint:
    """Read-only mapping ``HF tensor name -> tensor`` over a directory of safetensors files; a tensor is read from
    disk when it is asked for (``safe_open(...).get_tensor``), so a stage touches only its own layer range."""

    def __init__(self, path: str):
        self.path = path
        idx = os.path.join(path, INDEX)
        if os.path.exists(idx):
            with open(idx) as f:
                self.where: Dict[str, str] = dict(json.load(f)["weight_map"])
        else:
            from safetensors import safe_open
            files = [SINGLE] if os.path.exists(os.path.join(path, SINGLE)) else \
                sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
            if not files:
                raise FileNotFoundError(f"{path}: no safetensors files")
            self.where = {}
            for fn in files:
                with safe_open(os.path.join(path, fn), framework="pt") as f:
                    for k in f.keys():
                        self.where[k] = fn
        self.bytes_read = 0
        self._open: Dict[str, object] = {}

    def __contains__(self, name: str) -> bool:
        return name in self.where

    def keys(self) -> Iterator[str]:
        return iter(self.where)

    def __getitem__(self, name: str) -> torch.Tensor:
        from safetensors import safe_open
        fn = self.where[name]                       # KeyError names the missing tensor
        h = self._open.get(fn)
        if h is None:
            h = self._open[fn] = safe_open(os.path.join(self.path, fn), framework="pt")
        t = h.get_tensor(name)
        self.bytes_read += t.numel() * t.element_size()
        return t


def save_checkpoint(dm, path: str, link=None) -> None:
    """Every rank writes its stage's tensors; rank 0 adds ``config.json`` and the index over all ranks' files."""
    from safetensors.torch import save_file
    link = link or dm.link
    os.makedirs(path, exist_ok=True)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in dm.stage.params.hf_state_dict().items()}
    if dm.cfg.tied and "lm_head.weight" in sd and "model.embed_tokens.weight" in sd:
        sd.pop("lm_head.weight")                    # tied: stored once, like HF
    fn = f"model-{link.rank + 1:05d}-of-{link.world:05d}.safetensors"
    save_file(sd, os.path.join(path, fn), metadata={"format": "pt"})
    maps = link.all_gather_object({k: fn for k in sd})
    sizes = link.all_gather_object(sum(v.numel() * v.element_size() for v in sd.values()))
    if link.rank == 0:
        weight_map: Dict[str, str] = {}
        for m in maps:
            for k, f in m.items():
                weight_map.setdefault(k, f)         # a tied head on the last rank does not shadow the embedding
        with open(os.path.join(path, INDEX), "w") as f:
            json.dump({"metadata": {"total_size": int(sum(sizes))}, "weight_map": weight_map}, f, indent=1)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(config_to_json(dm.cfg), f, indent=1)
    link.barrier()
