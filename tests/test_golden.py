"""Oracle vs golden vectors produced by the reference's own LayerGroupModule + wire codec
(oracle/gen_golden.py; fixtures committed under tests/golden/)."""
import glob
import os

import pytest
import torch

from oracle import shard_oracle as O
from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.weights import init_state_dict

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_layergroup_*.pt")))


def test_fixtures_present():
    assert len(GOLDEN) == 6


def _oracle_hops(g, attn_mode):
    cfg = C.get_config(g["cfg"])
    sd = init_state_dict(cfg, seed=g["seed"])
    m = O.OracleModel(cfg, sd, attn_mode)
    ids = g["input_ids"]
    B, S = ids.shape
    with torch.no_grad():
        x = torch.nn.functional.embedding(ids, m.embed)
        cos, sin = O.rope_tables(cfg, torch.arange(S)[None].expand(B, -1), x.dtype)
        hops = []
        for a, b in g["bounds"]:
            x = O.wire_hop(O.shard_forward(cfg, m.layers[a:b], list(range(a, b)), x, cos, sin, attn_mode))
            hops.append(x)
        logits = torch.nn.functional.linear(O.rmsnorm(x, m.norm, cfg.rms_eps), m.head)[:, -4:, :]
    return hops, logits


@pytest.mark.parametrize("path", [p for p in GOLDEN if p.endswith("_eager.pt")], ids=os.path.basename)
def test_oracle_bit_exact_vs_reference_layergroup_eager(path):
    g = torch.load(path)
    hops, logits = _oracle_hops(g, "eager")
    for got, ref in zip(hops, g["hops"]):
        assert torch.equal(got, ref)
    assert torch.equal(logits, g["logits"])


@pytest.mark.parametrize("path", [p for p in GOLDEN if p.endswith("_sdpa.pt")], ids=os.path.basename)
def test_oracle_sdpa_math_vs_reference_layergroup_sdpa(path):
    """Tolerance: the spread between the reference's own eager and sdpa runs (its bf16 noise floor)."""
    g = torch.load(path)
    ge = torch.load(path.replace("_sdpa.pt", "_eager.pt"))
    hops, logits = _oracle_hops(g, "sdpa_math")
    for got, ref, ref_e in zip(hops, g["hops"], ge["hops"]):
        floor = O.rel_l2(ref_e, ref)
        assert O.rel_l2(got, ref) <= 1.25 * floor + 1e-6
    assert O.rel_l2(logits, g["logits"]) <= 1.25 * O.rel_l2(ge["logits"], g["logits"])
