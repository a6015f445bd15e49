"""CPU restatement of the reference's inter-shard WIRE path (rows a4/a5 of SURVEY.md §8).  TEST INFRASTRUCTURE:
only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this; the product never does.

What the reference does for every shard call, each way (tensorlink/ml/module.py:1536-1595 -> ml/worker.py:297-357):

  frame  = tensor_to_bytes(payload)            tensorlink/ml/utils.py:569-619
  (size, name) = store_in_shared_memory(frame) tensorlink/nodes/shared_memory.py:23-38     } once per process hop
  frame  = get_from_shared_memory(size, name)  tensorlink/nodes/shared_memory.py:6-20      } (user -> node -> node -> worker)
  payload = bytes_to_tensor(frame)             tensorlink/ml/utils.py:622-660

Frame layout (utils.py:611-619): ``[4-byte big-endian length of the JSON skeleton][JSON skeleton][safetensors blob]``.
The skeleton is the payload with every tensor replaced by ``{"__tensor_ref__": "__tensor_<i>__", "dtype": str(dtype),
"shape": [...]}`` (numbered in traversal order), tuples by ``{"__tuple__": true, "data": [...]}``, scalars kept,
anything else ``null``; the blob is ``safetensors.torch.save`` of ``{"__tensor_<i>__": tensor.cpu().contiguous()}``.
safetensors is a third-party dependency of the reference (pinned 0.4.5 in its lock file; 0.7.0 is installed here —
the blob format, an 8-byte header length + JSON header + raw little-endian data, is the same).
A ``DynamicCache`` becomes ``{"__dynamic_cache__": true, "key_cache": [...], "value_cache": [...]}`` (:599-605; the
reference reads transformers-4.x attribute names, so the golden frame is produced from a stand-in object carrying them);
``decode`` here keeps that dict instead of rebuilding an HF cache object.

Pinned by tests/test_wire_oracle.py against frames produced by the reference's own ``tensor_to_bytes``
(oracle/gen_golden_wire.py -> tests/golden/ref_wire_frames.pt): byte-identical.
"""
from __future__ import annotations

import json
import time
from multiprocessing import shared_memory
from typing import Any, Dict, Tuple

import torch
from safetensors.torch import load as _st_load
from safetensors.torch import save as _st_save

_SCALARS = (int, float, bool, str, type(None))


def encode(payload: Any) -> bytes:
    """utils.py:569-619."""
    tensors: Dict[str, torch.Tensor] = {}

    def skeleton(o):
        if isinstance(o, torch.Tensor):
            key = f"__tensor_{len(tensors)}__"
            tensors[key] = o.detach().cpu().contiguous().clone()            # the reference's extra host copy (:583)
            return {"__tensor_ref__": key, "dtype": str(o.dtype), "shape": list(o.shape)}
        if isinstance(o, dict):
            return {k: skeleton(v) for k, v in o.items()}
        if isinstance(o, tuple):
            return {"__tuple__": True, "data": [skeleton(v) for v in o]}
        if isinstance(o, list):
            return [skeleton(v) for v in o]
        if isinstance(o, _SCALARS):
            return o
        if o.__class__.__name__ == "DynamicCache":                           # :599-605 (transformers 4.x attribute names)
            return {"__dynamic_cache__": True, "key_cache": skeleton(o.key_cache), "value_cache": skeleton(o.value_cache)}
        return None                                                          # unserialisable objects are dropped (:607)

    head = json.dumps(skeleton(payload)).encode("utf-8")
    blob = _st_save(tensors) if tensors else b""
    return len(head).to_bytes(4, "big") + head + blob


def decode(frame: bytes) -> Any:
    """utils.py:622-660."""
    n = int.from_bytes(frame[:4], "big")
    skel = json.loads(frame[4:4 + n].decode("utf-8"))
    blob = frame[4 + n:]
    tensors = _st_load(blob) if blob else {}

    def restore(o):
        if isinstance(o, dict):
            if "__tensor_ref__" in o:
                return tensors[o["__tensor_ref__"]].to(dtype=getattr(torch, o["dtype"].replace("torch.", "")))
            if o.get("__tuple__"):
                return tuple(restore(v) for v in o["data"])
            return {k: restore(v) for k, v in o.items()}
        if isinstance(o, list):
            return [restore(v) for v in o]
        return o

    return restore(skel)


def shm_put(frame: bytes) -> Tuple[int, str]:
    """shared_memory.py:23-38 (encoded=True): create a POSIX segment of exactly len(frame) bytes, copy in."""
    seg = shared_memory.SharedMemory(create=True, size=max(len(frame), 1))
    seg.buf[:len(frame)] = frame
    name = seg.name
    seg.close()
    return len(frame), name


def shm_get(size: int, name: str) -> bytes:
    """shared_memory.py:6-20 (encoded=True): copy out (``tobytes``), copy again (``deepcopy``), unlink."""
    seg = shared_memory.SharedMemory(name=name)
    data = bytes(seg.buf[:size])
    again = bytes(bytearray(data))              # the reference deep-copies the bytes object it just made
    seg.close()
    seg.unlink()
    return again


def reference_hop(payload: Any, process_hops: int = 3) -> Any:
    """One direction of one shard call: encode, ``process_hops`` shared-memory hand-overs (user -> its node process,
    node -> worker's node process [a TCP stream between them, not modelled], node -> worker), decode."""
    frame = encode(payload)
    for _ in range(process_hops):
        frame = shm_get(*shm_put(frame))
    return decode(frame)


def time_reference_hop(payload: Any, repeats: int = 20, process_hops: int = 3) -> float:
    """Median seconds of ``reference_hop`` (host CPU; sockets, the reference's 0.1 s sleeps and polling not included)."""
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        reference_hop(payload, process_hops)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2]
