"""Generate golden wire frames with the reference's OWN codec (tensorlink/ml/utils.py:569-660).  TEST INFRA.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden_wire
Writes tests/golden/ref_wire_frames.pt: seeded payloads and the exact bytes ``tensor_to_bytes`` produced for them.
"""
import os

import torch

from oracle.ref_shim import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_wire_frames.pt")


class DynamicCache:
    """Stand-in with the transformers-4.x attribute names the reference codec reads (utils.py:600-603); the installed
    transformers 5.x class no longer has them.  Only the class NAME and these two lists matter."""

    def __init__(self, key_cache, value_cache):
        self.key_cache, self.value_cache = key_cache, value_cache


def cached_decode_payload():
    """What a reference user ships for one cached decode step of a 2-layer shard at position 7 (injector.py:508-556)."""
    g = torch.Generator().manual_seed(78)
    ks = [torch.randn(1, 2, 7, 16, generator=g).bfloat16() for _ in range(2)]
    vs = [torch.randn(1, 2, 7, 16, generator=g).bfloat16() for _ in range(2)]
    return {"hidden_states": torch.randn(1, 1, 64, generator=g).bfloat16(), "cache_position": torch.tensor([7]),
            "position_ids": torch.tensor([[7]]), "use_cache": True, "past_key_values": DynamicCache(ks, vs)}


def payloads():
    g = torch.Generator().manual_seed(77)
    hs = torch.randn(2, 5, 64, generator=g).bfloat16()
    cos, sin = torch.randn(2, 5, 16, generator=g).bfloat16(), torch.randn(2, 5, 16, generator=g).bfloat16()
    return {
        "decode_row": {"hidden_states": torch.randn(1, 1, 896, generator=g).bfloat16()},
        "live_ins": {"hidden_states": hs, "position_ids": torch.arange(5)[None].expand(2, -1).contiguous(),
                     "position_embeddings": (cos, sin), "use_cache": False, "past_key_values": None,
                     "causal_mask": torch.zeros(2, 1, 5, 5).bfloat16(), "kwargs": {}},
        "nested": [1, 2.5, "x", (torch.arange(6, dtype=torch.int64).view(2, 3), [torch.ones(3, dtype=torch.float32)])],
        "no_tensors": {"a": 1, "b": [True, None]},
        "dropped_object": {"keep": torch.zeros(2, dtype=torch.float16), "drop": object()},
    }


def reference_packets(frames):
    """FORWARD / BACKWARD packets built by the reference's own ``Torchnode.send_forward`` / ``send_backward``
    (p2p/torch_node.py:825-836, :865-869) and parsed back by its ``_handle_forward`` / ``_handle_backward`` (:251-299,
    :225-249), run as unbound functions on a recorder object (a real Torchnode needs sockets and RSA keys)."""
    import json
    import sys
    from oracle.ref_shim import REFERENCE_ROOT
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import tensorlink.p2p.torch_node as TN
    finally:
        sys.path.remove(REFERENCE_ROOT)
    mid = "ab" * 32

    class Rec:
        VERBOSE = 0
        role = "W"

        def __init__(self):
            self.sent, self.stored, self.nodes, self.modules = [], [], {"peer": None}, {mid: {"forward_queue": {}}}
            self.memory_manager = {}

        def send_to_node(self, node, data):
            self.sent.append(data)

        def debug_print(self, *a, **k):
            pass

        def _store_tensor_in_shared_memory(self, key, tensor, backward=False):
            self.stored.append((key, len(tensor), backward))

    class Peer:
        node_id, ghosts = "peer", 0

    rec, key = Rec(), [5, 1, mid]
    fwd_payload = len(frames["nested"]).to_bytes(8, "big") + frames["nested"] + frames["live_ins"]
    TN.Torchnode.send_forward(rec, None, fwd_payload, key, mid)
    TN.Torchnode.send_backward(rec, None, frames["decode_row"], key)
    fwd_packet, bwd_packet = rec.sent
    # the reference's own parsers on those packets
    eos = fwd_packet.find(b"::")
    size = int(fwd_packet[7:eos])
    tail = json.loads(fwd_packet[eos + 2 + size:])
    assert TN.Torchnode._handle_backward(rec, bwd_packet, Peer()) is True
    (bkey, bsize, bflag), = rec.stored
    return {"module_id": mid, "key": key, "forward_payload": fwd_payload, "backward_payload": frames["decode_row"],
            "forward_packet": fwd_packet, "backward_packet": bwd_packet,
            "ref_parsed_forward": {"size": size, "module_id": tail["module_id"], "key": tail["key"]},
            "ref_parsed_backward": {"size": bsize, "key": list(bkey)}}


def main():
    _, utils = import_reference()
    items = payloads()
    frames = {k: utils.tensor_to_bytes(v) for k, v in items.items()}
    for k, f in frames.items():                       # the reference decodes its own frames back
        back = utils.bytes_to_tensor(f)
        assert type(back) is type(items[k]) or items[k] is None
    items["dropped_object"]["drop"] = None            # what survives the reference codec (utils.py:607)
    cd = cached_decode_payload()
    frames_cache = utils.tensor_to_bytes(cd)
    pkv = cd["past_key_values"]                       # stored as plain lists; the tests rebuild the stand-in object
    cd["past_key_values"] = {"__dynamic_cache__": True, "key_cache": pkv.key_cache, "value_cache": pkv.value_cache}
    torch.save({"payloads": items, "frames": frames, "cached_decode": {"payload": cd, "frame": frames_cache},
                "packets": reference_packets(frames)}, OUT)
    print("wrote", OUT, {k: len(v) for k, v in frames.items()})


if __name__ == "__main__":
    main()
