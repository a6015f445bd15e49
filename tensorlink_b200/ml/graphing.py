"""Shard plan in the reference's schema.

The reference's ``ModelParser.create_distributed_config`` (/root/reference/tensorlink/ml/graphing.py:238-451) walks
an HF module tree and assigns sub-modules to workers memory-greedily, grouping consecutive decoder layers into
``offloaded_group`` entries (:20-61, :64-128).  On one NVSwitch box the plan is computed locally: contiguous layer
ranges per rank, embedding on the first rank, final norm + lm_head on the last.  The dict uses the same keys the
reference's hot path consumes (``type``, ``assigned_workers``, ``layer_range``, ``layer_paths``, ``num_layers``,
``memory``, ``module``, ``training``, ``optimizer_type``, ``parent_module_path``; graphing.py:44-55), so a plan
produced by the reference's planner is accepted by ``DistributedModel(config=...)`` as well.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

from .configs import ShardModelConfig


def split_even(n_layers: int, n_stages: int) -> List[range]:
    base, rem = divmod(n_layers, n_stages)
    out, a = [], 0
    for i in range(n_stages):
        b = a + base + (1 if i < rem else 0)
        out.append(range(a, b))
        a = b
    return out


def split_balanced(cfg: ShardModelConfig, n_stages: int) -> List[range]:
    """Byte-balanced split: lm_head (2·H·V bytes) counts as head_cost layers on the last stage and the embedding
    gather is free, so the last stage gets fewer layers (SURVEY.md §7.2 'stage balance')."""
    if n_stages == 1:
        return [range(cfg.n_layers)]
    head_cost = cfg.vocab * cfg.hidden / cfg.layer_params()
    per = (cfg.n_layers + head_cost) / n_stages
    last = max(1, min(cfg.n_layers - (n_stages - 1), round(per - head_cost)))
    rest = split_even(cfg.n_layers - last, n_stages - 1)
    return rest + [range(cfg.n_layers - last, cfg.n_layers)]


def make_plan(cfg: ShardModelConfig, n_stages: int, training: bool = False, balanced: bool = False,
              optimizer_type: str = "adam") -> Dict[str, dict]:
    ranges = split_balanced(cfg, n_stages) if balanced else split_even(cfg.n_layers, n_stages)
    layer_bytes = 2 * cfg.layer_params()
    plan: Dict[str, dict] = {
        "model.embed_tokens": {"type": "loaded", "name": cfg.name, "assigned_workers": [0],
                               "memory": 2 * cfg.vocab * cfg.hidden, "module": "Embedding", "training": training,
                               "parent_module_path": "model"},
    }
    for rank, r in enumerate(ranges):
        a, b = r.start, r.stop - 1
        plan[f"model.layers.{a}-{b}"] = {
            "type": "offloaded_group", "name": cfg.name, "assigned_workers": [rank], "layer_range": (a, b),
            "layer_paths": [f"model.layers.{i}" for i in r], "memory": layer_bytes * len(r),
            "module": "Qwen3DecoderLayer" if cfg.qk_norm else "Qwen2DecoderLayer", "training": training,
            "optimizer_type": optimizer_type, "num_layers": len(r), "parent_module_path": "model"}
    last = n_stages - 1
    plan["model.norm"] = {"type": "loaded", "name": cfg.name, "assigned_workers": [last], "memory": 2 * cfg.hidden,
                          "module": "RMSNorm", "training": training, "parent_module_path": "model"}
    plan["lm_head"] = {"type": "loaded", "name": cfg.name, "assigned_workers": [last],
                       "memory": 2 * cfg.vocab * cfg.hidden, "module": "Linear", "training": training,
                       "tied_to": "model.embed_tokens" if cfg.tied else None, "parent_module_path": ""}
    return plan


def worker_ranks(plan: Dict[str, dict]) -> Dict[object, int]:
    """Worker id -> pipeline rank.  Our own plans name ranks directly (ints).  The reference's planner names workers
    by their node-id hash strings (graphing.py:730-761 `_try_assign_worker` appends the worker's id); those are
    mapped to ranks in pipeline order = the order in which they first appear walking the plan by layer position."""
    ids = []
    for key, e in sorted(plan.items(), key=lambda kv: _plan_position(kv[0], kv[1])):
        for w in e.get("assigned_workers", []):
            if w not in ids:
                ids.append(w)
    if all(isinstance(w, int) or (isinstance(w, str) and w.isdigit()) for w in ids):
        return {w: int(w) for w in ids}
    return {w: i for i, w in enumerate(ids)}


def _plan_position(key: str, e: dict):
    """Sort key that walks a plan front to back: embedding, decoder layers by index, final norm, lm_head."""
    if "layer_range" in e:
        return (1, e["layer_range"][0])
    if ".layers." in key and key.rsplit(".", 1)[1].isdigit():
        return (1, int(key.rsplit(".", 1)[1]))
    if "embed" in key:
        return (0, 0)
    if key.endswith("lm_head") or key == "lm_head":
        return (3, 0)
    if key.endswith("norm"):
        return (2, 0)
    return (1, 1 << 30)


def stage_layers(plan: Dict[str, dict], rank: int) -> List[int]:
    """Layer ids assigned to ``rank`` by a plan in the reference schema."""
    out: List[int] = []
    wr = worker_ranks(plan)
    _ranks = lambda e: [wr[w] for w in e.get("assigned_workers", [])]   # noqa: E731
    for key, e in plan.items():
        if e.get("type") == "offloaded_group" and rank in _ranks(e):
            a, b = e["layer_range"]
            out += list(range(a, b + 1))
        elif e.get("type") == "offloaded" and rank in _ranks(e) and ".layers." in key:
            out.append(int(key.rsplit(".", 1)[1]))
    return sorted(out)


def n_stages(plan: Dict[str, dict]) -> int:
    return 1 + max(worker_ranks(plan).values(), default=0)
