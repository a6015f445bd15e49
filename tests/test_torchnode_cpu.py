"""Packet adapter (tensorlink_b200/p2p/torch_node.py) against the reference's own packet builders / parsers.

The golden packets were produced by the reference's ``Torchnode.send_forward`` / ``send_backward`` bodies
(p2p/torch_node.py:825-836, :865-869) in oracle/gen_golden_wire.py; here the product builders must reproduce them byte
for byte, the parsers must recover what the reference's ``_handle_forward`` / ``_handle_backward`` recover, and a fake
worker shows the request -> reply flow (the GPU flow is tests/test_worker_gpu.py)."""
import os
import pickle

import pytest
import torch

from tensorlink_b200.p2p import torch_node as T
from tensorlink_b200.p2p import wire

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wire_frames.pt")


def test_packets_equal_the_reference_builders():
    g = torch.load(GOLD)["packets"]
    mid = g["module_id"]
    assert T.build_forward(g["forward_payload"], g["key"], mid) == g["forward_packet"]
    assert T.build_backward(g["backward_payload"], g["key"]) == g["backward_packet"]
    payload, module_id, key = T.parse_forward(g["forward_packet"])
    assert payload == g["forward_payload"] and module_id == mid and key == tuple(g["key"])
    payload, tag = T.parse_backward(g["backward_packet"])
    assert payload == g["backward_payload"] and tag == tuple(g["key"])
    # what the reference's own parser recovered from the same packets (recorded by the generator)
    assert g["ref_parsed_forward"] == {"size": len(g["forward_payload"]), "module_id": mid, "key": g["key"]}
    assert g["ref_parsed_backward"] == {"size": len(g["backward_payload"]), "key": g["key"]}


class _EchoWorker:
    """Stands in for DistributedWorker: doubles hidden_states / the gradient (no GPU in this test)."""

    def __init__(self, mid):
        self.modules = {mid: object()}
        self.calls = []

    def handle_forward_frame(self, module_id, key, data):
        _args, kwargs = wire.unpack_forward(data)
        self.calls.append(("fwd", module_id, key))
        return wire.encode(dict(kwargs, hidden_states=kwargs["hidden_states"] * 2))

    def handle_backward_frame(self, module_id, tag, data):
        self.calls.append(("bwd", module_id, tag))
        return wire.encode(wire.decode(data) * 2)


def test_request_reply_flow_and_errors():
    g = torch.load(GOLD)["packets"]
    mid = g["module_id"]
    node = T.B200Torchnode(_EchoWorker(mid))
    x = torch.arange(6, dtype=torch.float32).view(1, 2, 3).bfloat16()
    req = T.build_forward(wire.pack_forward((), {"hidden_states": x, "use_cache": False}), [3, 0, mid], mid)
    reply = node.handle_data(req + T.EOT)                              # the EOT marker of Connection.send is tolerated
    payload, module_id, key = T.parse_forward(reply)
    assert module_id == mid and key == (3, 0, mid)
    out = wire.decode(pickle.loads(payload))                           # the user side unpickles (shared_memory.py:10-11)
    assert torch.equal(out["hidden_states"], x * 2) and out["use_cache"] is False
    grad = torch.ones(1, 2, 3).bfloat16()
    back = node.handle_data(T.build_backward(wire.encode(grad), [3, 0, mid]))
    payload, tag = T.parse_backward(back)
    assert tag == (3, 0, mid) and torch.equal(wire.decode(payload), grad * 2)
    assert node.worker.calls == [("fwd", mid, (3, 0, mid)), ("bwd", mid, (3, 0, mid))]
    assert node.handle_data(b"OPTIMIZER{}") is None                    # other packet types are not this adapter's business
    with pytest.raises(KeyError):
        node.handle_data(T.build_forward(b"x", [0, 0, "f" * 64], "f" * 64))
    assert node.ghosts == 1
    with pytest.raises(ValueError):
        T.parse_forward(req[:-40][:30])
