"""Product-side wire codec (tensorlink_b200/p2p/wire.py) against frames produced by the reference's own codec and
against the oracle restatement: byte-identical frames, exact round trips, loud errors on truncated input.  CPU only."""
import os

import pytest
import torch

from oracle import wire_oracle as W
from tensorlink_b200.p2p import wire

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wire_frames.pt")


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b and type(a) is type(b)


def test_frames_equal_the_reference_codec_and_the_oracle():
    g = torch.load(GOLD)
    for name, payload in g["payloads"].items():
        assert wire.encode(payload) == g["frames"][name] == W.encode(payload), name
        assert _same(wire.decode(g["frames"][name]), payload), name


def test_forward_request_roundtrip_in_the_reference_layout():
    g = torch.load(GOLD)
    args, kwargs = (g["payloads"]["decode_row"]["hidden_states"],), g["payloads"]["live_ins"]
    data = wire.pack_forward(args, kwargs)
    # layout the reference worker parses (ml/worker.py:303-307): 8-byte length, args frame, kwargs frame
    n = int.from_bytes(data[:8], "big")
    assert data[8:8 + n] == W.encode(args) and data[8 + n:] == W.encode(kwargs)
    a2, k2 = wire.unpack_forward(data)
    assert _same(a2, args) and _same(k2, kwargs)


def test_truncated_input_is_rejected():
    g = torch.load(GOLD)
    f = g["frames"]["live_ins"]
    for bad in (b"", f[:3], f[:40]):
        with pytest.raises(ValueError):
            wire.decode(bad)
    with pytest.raises(ValueError):
        wire.unpack_forward(bytes(4))
    with pytest.raises(ValueError):
        wire.unpack_forward((10 ** 6).to_bytes(8, "big") + f)


class DynamicCache:
    """Same stand-in as oracle/gen_golden_wire.py: the class name and the 4.x attribute names are what the codec keys on."""

    def __init__(self, key_cache, value_cache):
        self.key_cache, self.value_cache = key_cache, value_cache


def test_dynamic_cache_branch_matches_the_reference_codec():
    """utils.py:599-605: a cached decode step ships the whole DynamicCache; the frame the reference's own codec produced
    for it (golden) is reproduced byte for byte by the product codec and by the oracle, from the object form."""
    g = torch.load(GOLD)["cached_decode"]
    payload = dict(g["payload"])
    as_dict = payload["past_key_values"]
    payload["past_key_values"] = DynamicCache(as_dict["key_cache"], as_dict["value_cache"])
    assert wire.encode(payload) == g["frame"] == W.encode(payload)
    back = wire.decode(g["frame"])
    assert _same(back, g["payload"])                      # decoded form: the dict (re-encodes to the same bytes)
    assert wire.encode(back) == g["frame"]


def test_dynamic_cache_with_transformers5_layout():
    """transformers 5.x keeps keys/values in ``cache.layers[i].keys/.values``; the product codec reads both layouts."""
    from transformers import DynamicCache as HFCache
    g = torch.load(GOLD)["cached_decode"]["payload"]["past_key_values"]
    c = HFCache()
    for li, (k, v) in enumerate(zip(g["key_cache"], g["value_cache"])):
        c.update(k, v, li)
    frame = wire.encode({"past_key_values": c})
    back = wire.decode(frame)["past_key_values"]
    assert back["__dynamic_cache__"] is True and _same(back["key_cache"], g["key_cache"]) and _same(back["value_cache"], g["value_cache"])
