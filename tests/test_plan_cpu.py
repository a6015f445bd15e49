"""Host-side logic that needs no GPU: shard plan schema (ml/graphing.py), parameter arena layout (ml/shard.py),
node shims, seeded weights."""
import pytest
import torch

from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml import graphing
from tensorlink_b200.ml.shard import ShardParams
from tensorlink_b200.ml.weights import init_state_dict, synthetic_tokens


@pytest.mark.parametrize("cfg,n", [(C.QWEN25_7B, 1), (C.QWEN25_7B, 4), (C.QWEN25_7B, 8), (C.QWEN3_8B, 8), (C.QWEN25_05B, 3)])
def test_plan_schema_and_coverage(cfg, n):
    for balanced in (False, True):
        plan = graphing.make_plan(cfg, n, training=True, balanced=balanced)
        assert graphing.n_stages(plan) == n
        seen = []
        for rank in range(n):
            seen += graphing.stage_layers(plan, rank)
        assert seen == list(range(cfg.n_layers))                      # every layer exactly once, in order
        groups = [e for e in plan.values() if e["type"] == "offloaded_group"]
        for e in groups:                                              # the reference's keys (ml/graphing.py:44-55)
            for k in ("type", "name", "assigned_workers", "layer_range", "layer_paths", "memory", "module", "training",
                      "optimizer_type", "num_layers", "parent_module_path"):
                assert k in e
            a, b = e["layer_range"]
            assert e["num_layers"] == b - a + 1 == len(e["layer_paths"])
        assert plan["model.embed_tokens"]["assigned_workers"] == [0]
        assert plan["lm_head"]["assigned_workers"] == [n - 1]
        assert (plan["lm_head"]["tied_to"] == "model.embed_tokens") == cfg.tied


def test_balanced_split_gives_the_head_stage_fewer_layers():
    r = graphing.split_balanced(C.QWEN25_7B, 8)
    assert [len(x) for x in r] == [4, 4, 4, 4, 4, 4, 3, 1]            # lm_head ~ 2.3 layers of bytes
    assert [len(x) for x in graphing.split_even(28, 8)] == [4, 4, 4, 4, 3, 3, 3, 3]


def test_plan_from_reference_style_dict_is_accepted():
    """A plan written the way the reference's ModelParser emits it (string worker ids, 'offloaded' single layers)."""
    plan = {
        "model.layers.0-1": {"type": "offloaded_group", "assigned_workers": ["0"], "layer_range": (0, 1)},
        "model.layers.2": {"type": "offloaded", "assigned_workers": ["1"]},
        "model.layers.3": {"type": "offloaded", "assigned_workers": ["1"]},
    }
    assert graphing.stage_layers(plan, 0) == [0, 1] and graphing.stage_layers(plan, 1) == [2, 3]
    assert graphing.n_stages(plan) == 2


@pytest.mark.parametrize("cfg", [C.TINY_QWEN2, C.TINY_QWEN3], ids=lambda c: c.name)
def test_param_arena_round_trip(cfg):
    """fused-QKV / interleaved gate-up arena <-> HF state dict is lossless; every view is 256-byte aligned."""
    sd = init_state_dict(cfg)
    p = ShardParams(cfg, range(cfg.n_layers), True, True, "cpu")
    p.load_hf_state_dict(sd)
    back = p.hf_state_dict()
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    for name, (off, n, shape) in p.offsets.items():
        assert (off * 2) % 256 == 0
    wgu = p.v[f"l0.wgu"]
    assert torch.equal(wgu[0::2], sd["model.layers.0.mlp.gate_proj.weight"])
    assert torch.equal(wgu[1::2], sd["model.layers.0.mlp.up_proj.weight"])
    if cfg.tied:
        assert p.v["head"].data_ptr() == p.v["embed"].data_ptr()


def test_seeded_weights_are_per_tensor_reproducible():
    cfg = C.TINY_QWEN2
    full = init_state_dict(cfg)
    part = init_state_dict(cfg, layers=[2], with_embed=False, with_head=False)
    for k, v in part.items():
        assert torch.equal(v, full[k])
    a, b = synthetic_tokens(cfg, 2, 5), synthetic_tokens(cfg, 2, 5)
    assert torch.equal(a, b) and a.dtype == torch.int64 and int(a.max()) < cfg.vocab


def test_node_shim_contract():
    from tensorlink_b200.nodes.nodes import User, UserConfig
    u = User(config=UserConfig())
    assert u.__class__.__name__ == "User"                              # triggers auto-distribution (module.py:345-346)
    for attr in ("node_requests", "node_responses", "mpc_lock", "send_request"):
        assert hasattr(u, attr)
    with pytest.raises(NotImplementedError):
        u.send_request("request_job", None)


def test_apply_eos_matches_hf_generate_stopping():
    """Host logic of generate(eos_token_id=..., pad_token_id=...): applied to a full greedy generation it must give
    what HF's own stopping criteria give (rows padded after their first EOS, output ends when every row is done)."""
    import torch
    from tensorlink_b200.ml import configs as C
    from tensorlink_b200.ml.module import apply_eos
    from tensorlink_b200.ml.weights import init_state_dict, synthetic_tokens
    from tests.hf_util import hf_model
    cfg = C.TINY_QWEN2
    hf = hf_model(cfg, init_state_dict(cfg, dtype=torch.float32), "eager", torch.float32)
    ids = synthetic_tokens(cfg, 3, 6)
    with torch.no_grad():
        full = hf.generate(ids, max_new_tokens=12, do_sample=False, eos_token_id=None, pad_token_id=0)
        for eos in (int(full[0, 8]), [int(full[1, 7]), int(full[2, 15])], int(full[0, 17]), cfg.vocab - 1):
            want = hf.generate(ids, max_new_tokens=12, do_sample=False, eos_token_id=eos, pad_token_id=0)
            got = apply_eos(full, 6, eos, 0)
            assert got.shape == want.shape and torch.equal(got, want), eos
    assert torch.equal(apply_eos(full, 6), full)


# ---------------------------------------------------------------------------------------------- the reference's planner
def _golden_plans():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_plans.json")) as f:
        return json.load(f)


def _norm(x):
    """JSON turns tuples into lists; compare structurally."""
    if isinstance(x, (list, tuple)):
        return [_norm(v) for v in x]
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    return x


def test_planner_restatement_equals_the_reference_planner():
    """``graphing.create_distributed_config`` (memory estimate, greedy assignment, grouping) against plans produced by the
    reference's own ``ModelParser`` (tests/golden/ref_plans.json, oracle/gen_golden_plans.py): same keys in the same
    order, same workers, same layer ranges, memory to the last digit."""
    import pytest
    from tensorlink_b200.ml import graphing
    from tensorlink_b200.ml.configs import get_config
    for tag, g in _golden_plans().items():
        got = graphing.create_distributed_config(get_config(g["model"]), {w: {"gpu_memory": m} for w, m in g["workers"].items()},
                                                 trusted=False, **g["kwargs"])
        assert got["success"], tag
        assert got["model_memory"] == pytest.approx(g["model_memory"], rel=1e-12), tag
        assert list(got["config"]) == list(g["config"]), tag
        for k, want in g["config"].items():
            have = _norm(got["config"][k])
            assert set(have) == set(want), (tag, k, set(have) ^ set(want))
            for f, v in want.items():
                if f == "memory":
                    assert have[f] == pytest.approx(v, rel=1e-12), (tag, k)
                else:
                    assert have[f] == v, (tag, k, f, have[f], v)


def test_reference_plans_build_the_same_stages():
    """A plan produced by the reference's planner (workers named by node-id hash strings, lm_head on a worker of its
    own) is consumed as-is: workers map to pipeline ranks in plan order, every rank gets exactly the plan's layers."""
    from tensorlink_b200.ml import graphing
    for tag, g in _golden_plans().items():
        plan = g["config"]
        wr = graphing.worker_ranks(plan)
        assert sorted(wr.values()) == list(range(len(wr))), tag
        n = graphing.n_stages(plan)
        covered = []
        for rank in range(n):
            layers = graphing.stage_layers(plan, rank)
            assert layers == sorted(layers) and (not layers or layers == list(range(layers[0], layers[-1] + 1))), (tag, rank)
            covered += layers
            want = [i for k, e in plan.items() if e.get("type") == "offloaded_group" and wr[e["assigned_workers"][0]] == rank
                    for i in range(e["layer_range"][0], e["layer_range"][1] + 1)]
            assert layers == want, (tag, rank)
        from tensorlink_b200.ml.configs import get_config
        assert covered == list(range(get_config(g["model"]).n_layers)), tag
    hashed = _golden_plans()["qwen25_7b_infer_hashed_workers"]["config"]
    assert graphing.n_stages(hashed) == 2 and graphing.stage_layers(hashed, 0) == list(range(0, 14))


def test_balanced_mode_and_failure():
    from tensorlink_b200.ml import graphing
    from tensorlink_b200.ml.configs import QWEN25_7B
    workers = {f"w{i}": {"gpu_memory": 40e9} for i in range(4)}
    bal = graphing.create_distributed_config(QWEN25_7B, {w: {"gpu_memory": 6e9} for w in workers}, training=False, max_seq_len=1024,
                                             balanced=True)
    sizes = [len(graphing.stage_layers(bal["config"], r)) for r in range(4)]
    assert sizes == [len(r) for r in graphing.split_balanced(QWEN25_7B, 4)] and sum(sizes) == 28
    tight = graphing.create_distributed_config(QWEN25_7B, {"a": {"gpu_memory": 2e9}}, training=False, max_seq_len=1024)
    assert tight["success"] is False                        # like the reference: partial config, success False
