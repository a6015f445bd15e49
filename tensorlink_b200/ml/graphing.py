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

from typing import Dict, List

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


# ---------------------------------------------------------------------------------------------------------------------
# The reference's planner, restated over the fixed module tree of a Qwen2/Qwen3 causal LM
#
#   model (ForCausalLM)                                  depth 0
#     model.model (decoder stack)                        depth 1   loop-iterable: never assigned whole (graphing.py:591-598)
#       model.model.embed_tokens | layers | norm | rotary_emb        depth 2   (layers = ModuleList: recursed into)
#         model.model.layers.<i>                         depth 3
#     model.lm_head                                      depth 1
#
# ``ModelParser.create_distributed_config`` (graphing.py:238-451) walks this tree depth-first, asks ``estimate_memory``
# (utils.py:36-124) for each module and hands it to the first worker with room, preferring the worker of the previous
# module (`_try_assign_worker`, :730-761); consecutive layers on one worker are then merged into ``offloaded_group``
# entries (`_group_sequential_layers`, :64-128).  tests/test_plan_cpu.py pins this restatement to plans the reference's
# own planner produced (tests/golden/ref_plans.json, generated by oracle/gen_golden_plans.py).

_OVERHEAD = 1.20          # utils.py:119
_DTYPE_SIZE = 2           # estimate_memory's default dtype is float16 (utils.py:41): activation / KV element size


def _heuristic_hidden(n_params: int) -> int:
    """utils.py:86-88: modules without a hidden-size attribute."""
    return max(256, min(int((n_params / 12) ** 0.5), 8192))


def estimate_memory(cfg: ShardModelConfig, module: str, training: bool = True, batch_size: int = 256, seq_length: int = 2048,
                    optimizer_type: str = "adam", include_kv_cache: bool = True, param_bytes_per_el: int = 2):
    """``estimate_memory`` (utils.py:36-124) for one module of the tree above: ``module`` in {"model", "model.model",
    "embed_tokens", "layers", "layer", "norm", "rotary_emb", "lm_head"}.  Returns (total bytes, breakdown)."""
    H, V, L = cfg.hidden, cfg.vocab, cfg.n_layers
    layer_p = cfg.layer_params()
    rot_buf = 2 * (cfg.head_dim // 2) * param_bytes_per_el   # inv_freq + original_inv_freq of the rotary module (cast with the skeleton)
    n_params = {"embed_tokens": V * H, "lm_head": V * H, "norm": H, "rotary_emb": 0, "layer": layer_p, "layers": L * layer_p,
                "model.model": V * H + L * layer_p + H,
                "model": V * H + L * layer_p + H + (0 if cfg.tied else V * H)}[module]
    buf_bytes = rot_buf if module in ("rotary_emb", "model.model", "model") else 0
    b = {"parameters": n_params * param_bytes_per_el + buf_bytes, "gradients": 0, "optimizer": 0, "activations": 0, "kv_cache": 0}
    if training:
        b["gradients"] = b["parameters"]
        b["optimizer"] = 2 * b["parameters"] * (4 / _DTYPE_SIZE) if optimizer_type.lower() in ("adam", "adamw") else b["parameters"]
    has_config = module in ("model", "model.model", "rotary_emb")       # HF modules that carry ``.config``
    hidden = H if (has_config or module == "layer") else _heuristic_hidden(n_params)   # decoder layers carry hidden_size
    b["activations"] = batch_size * seq_length * hidden * _DTYPE_SIZE * (7 if training else 4)
    if include_kv_cache and has_config and not training:
        head_dim = H // cfg.n_heads                                       # utils.py:104 (not config.head_dim)
        b["kv_cache"] = batch_size * seq_length * L * cfg.n_kv_heads * head_dim * 2 * _DTYPE_SIZE
    return sum(b.values()) * _OVERHEAD, b


class AssignmentError(Exception):
    """graphing.py:14-17."""


def create_distributed_config(cfg: ShardModelConfig, workers: Dict[str, dict], training: bool, trusted: bool = False,
                              optimizer_type: str = "adam", optimizer_spec: dict = None, host_max_memory_bytes: int = 0,
                              host_max_module_bytes: int = 0, host_max_depth: int = 2, max_offload_depth: int = 3,
                              max_seq_len: int = 4096, batch_size: int = 1, model_type: str = "chat", name: str = "",
                              balanced: bool = False) -> dict:
    """The reference's memory-greedy shard plan (same arguments, same result dict: success / config / model_memory /
    host_memory_used).  ``workers``: id -> {"gpu_memory": bytes}.  ``balanced=True`` instead splits the layers so every
    worker streams the same bytes per decode step, counting the lm_head on the last one (``split_balanced``) — the
    reference's greedy fill leaves the last workers short or idle, which costs pipeline throughput on equal GPUs."""
    optimizer_spec = optimizer_spec or {}
    state = {w: {"gpu_memory": float(i["gpu_memory"])} for w, i in workers.items()}
    assigned_host = [0.0]
    host_cap_module = host_max_module_bytes or 1e15
    layer_mod = "Qwen3DecoderLayer" if cfg.qk_norm else "Qwen2DecoderLayer"
    norm_mod = "Qwen3RMSNorm" if cfg.qk_norm else "Qwen2RMSNorm"
    rot_mod = "Qwen3RotaryEmbedding" if cfg.qk_norm else "Qwen2RotaryEmbedding"
    top_mod = "Qwen3ForCausalLM" if cfg.qk_norm else "Qwen2ForCausalLM"
    # the reference detects tying by ``data_ptr`` equality (:398-414), which holds for EVERY model on its meta-device
    # skeletons, so lm_head always carries ``tied_to`` and both ends are "force_host" candidates
    tied_embed, tied_head = "model.model.embed_tokens", "model.lm_head"
    config: Dict[str, dict] = {}

    def mem(kind, depth):
        total, b = estimate_memory(cfg, kind, training=training, seq_length=max_seq_len, optimizer_type=optimizer_type,
                                   batch_size=batch_size, include_kv_cache=(depth == 0))
        if depth > 0:                                  # children are recursed with count_activations=False (:689)
            total -= b["activations"]
        return total

    def try_assign(memory, last):
        order = [w for w in state if w == last] + [w for w in state if w != last]
        if len(order) > 1:
            order = order[:1] + sorted(order[1:], key=lambda w: state[w]["gpu_memory"], reverse=True)
        for w in order:
            if state[w]["gpu_memory"] >= memory:
                state[w]["gpu_memory"] -= memory
                return w
        return None

    def entry(path, kind, module, depth, last, parent):
        """One assignable (non-loop) module: host if allowed, else a worker.  Returns the worker (or None for host)."""
        memory = mem(kind, depth)
        force_host = path in (tied_embed, tied_head)
        if (host_max_memory_bytes and memory <= host_max_memory_bytes - assigned_host[0] and depth <= host_max_depth
                and memory <= host_cap_module) or force_host:
            if force_host and memory > host_max_memory_bytes - assigned_host[0]:
                pass                                   # "Don't force it if it truly won't fit" (:551-553)
            else:
                assigned_host[0] += memory
                config[path] = {"type": "loaded", "device": "host", "name": name, "memory": memory, "module": module,
                                "module_path": path, "training": training, "optimizer_spec": optimizer_spec,
                                "batch_size": batch_size, "model_type": model_type, "input_boundary": False}
                if path == tied_head:
                    config[path]["tied_to"] = tied_embed
                return None, True
        w = try_assign(memory, last)
        if w is None:
            return None, False
        config[path] = {"type": "offloaded", "name": name, "assigned_workers": [w], "memory": memory, "module": module,
                        "module_path": path, "training": training, "optimizer_spec": optimizer_spec, "batch_size": batch_size,
                        "model_type": model_type}
        if path == tied_head:
            config[path]["tied_to"] = tied_embed
        if parent is not None:
            config[path]["parent_module_path"] = parent
        return w, True

    model_memory, _ = estimate_memory(cfg, "model", training=training, seq_length=max_seq_len, optimizer_type=optimizer_type,
                                      batch_size=batch_size, include_kv_cache=True)
    success = True
    try:
        # depth 0: the whole model on one worker if it fits (never kept on the host: depth 0 > nothing, :541-546 applies
        # only with a host budget, which the memory of a whole LM never meets in practice)
        w, ok = entry("model", "model", top_mod, 0, None, None)
        if not ok:
            last = None
            # model.model is loop-iterable (:591-598) -> its children, in registration order
            w, ok = entry("model.model.embed_tokens", "embed_tokens", "Embedding", 2, last, "model.model")
            if not ok:
                raise AssignmentError("Unable to assign model.model.embed_tokens: no children to distribute")
            last = w or last
            if balanced:
                ws = list(state)
                for r, rng in enumerate(split_balanced(cfg, len(ws))):
                    for i in rng:
                        m = mem("layer", 3)
                        state[ws[r]]["gpu_memory"] -= m
                        config[f"model.model.layers.{i}"] = {
                            "type": "offloaded", "name": name, "assigned_workers": [ws[r]], "memory": m, "module": layer_mod,
                            "module_path": f"model.model.layers.{i}", "training": training, "optimizer_spec": optimizer_spec,
                            "batch_size": batch_size, "model_type": model_type, "parent_module_path": "model.model.layers"}
                last = ws[-1]
            else:
                for i in range(cfg.n_layers):
                    w, ok = entry(f"model.model.layers.{i}", "layer", layer_mod, 3, last, "model.model.layers")
                    if not ok:
                        config[f"model.model.layers.{i}"] = {"type": "unassigned", "required_memory": mem("layer", 3),
                                                             "module_path": f"model.model.layers.{i}",
                                                             "reason": f"Exceeded max recursion depth ({max_offload_depth})"}
                        raise AssignmentError(f"Unable to assign model.model.layers.{i}: no children to distribute")
                    last = w or last
            for path, kind, module in (("model.model.norm", "norm", norm_mod), ("model.model.rotary_emb", "rotary_emb", rot_mod)):
                w, ok = entry(path, kind, module, 2, last, "model.model")
                if not ok:
                    raise AssignmentError(f"Unable to assign {path}: no children to distribute")
                last = w or last
            w, ok = entry("model.lm_head", "lm_head", "Linear", 1, last, "model")
            if not ok:
                raise AssignmentError("Unable to assign model.lm_head: no children to distribute")
        config = _group_sequential_layers(config)
    except AssignmentError:
        success = False
    return {"success": success, "config": config, "model_memory": model_memory, "host_memory_used": assigned_host[0]}


def _group_sequential_layers(config: Dict[str, dict]) -> Dict[str, dict]:
    """graphing.py:64-128 + :20-61: consecutive layers on one worker become one ``offloaded_group`` entry (placed first,
    like the reference does), everything else follows in its original order."""
    import re
    runs: List[list] = []
    for path, e in config.items():
        m = re.match(r"^(.+\.)(\d+)$", path)
        if e.get("type") != "offloaded" or not m:
            continue
        idx, worker = int(m.group(2)), (e["assigned_workers"][0] if e["assigned_workers"] else None)
        if runs and runs[-1][0] == m.group(1) and runs[-1][1] == worker and runs[-1][2][-1][0] == idx - 1:
            runs[-1][2].append((idx, path, e))
        else:
            runs.append([m.group(1), worker, [(idx, path, e)]])
    out: Dict[str, dict] = {}
    done = set()
    for parent, worker, group in runs:
        done.update(p for _, p, _ in group)
        if len(group) == 1:
            out[group[0][1]] = group[0][2]
            continue
        first = group[0][2]
        g = {"type": "offloaded_group", "name": first.get("name", ""), "assigned_workers": [worker],
             "layer_range": (group[0][0], group[-1][0]), "layer_paths": [p for _, p, _ in group],
             "memory": sum(e.get("memory", 0) for _, _, e in group), "module": first.get("module", ""),
             "training": first.get("training", False), "optimizer_type": first.get("optimizer_type", "adam"),
             "num_layers": len(group)}
        if "parent_module_path" in first:
            g["parent_module_path"] = first["parent_module_path"]
        out[f"{parent}{group[0][0]}-{group[-1][0]}"] = g
    for path, e in config.items():
        if path not in done:
            out[path] = e
    return out
