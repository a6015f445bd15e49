"""The wire oracle (oracle/wire_oracle.py) against frames produced by the reference's own codec
(oracle/gen_golden_wire.py ran ``tensorlink.ml.utils.tensor_to_bytes`` unmodified): byte-identical encoding, exact
decoding, and the shared-memory hand-over is lossless.  CPU only."""
import os

import torch

from oracle import wire_oracle as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wire_frames.pt")


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b and type(a) is type(b)


def test_encode_is_byte_identical_to_the_reference_codec():
    g = torch.load(GOLD)
    assert set(g["frames"]) == {"decode_row", "live_ins", "nested", "no_tensors", "dropped_object"}
    for name, payload in g["payloads"].items():
        assert W.encode(payload) == g["frames"][name], name


def test_decode_restores_reference_frames_exactly():
    g = torch.load(GOLD)
    for name, frame in g["frames"].items():
        assert _same(W.decode(frame), g["payloads"][name]), name


def test_unserialisable_objects_become_none_like_the_reference():
    g = torch.load(GOLD)
    payload = dict(g["payloads"]["dropped_object"], drop=object())
    assert W.encode(payload) == g["frames"]["dropped_object"]


def test_shared_memory_handover_and_whole_hop_are_lossless():
    g = torch.load(GOLD)
    frame = g["frames"]["live_ins"]
    size, name = W.shm_put(frame)
    assert size == len(frame) and W.shm_get(size, name) == frame
    assert _same(W.reference_hop(g["payloads"]["live_ins"]), g["payloads"]["live_ins"])
    empty = W.encode({})
    assert W.shm_get(*W.shm_put(empty)) == empty
    assert W.time_reference_hop(g["payloads"]["decode_row"], repeats=3) > 0
