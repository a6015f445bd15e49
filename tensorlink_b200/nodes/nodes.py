"""Local stand-ins for the reference's role objects (/root/reference/tensorlink/nodes/nodes.py).

``DistributedModel(node=...)`` in the reference needs an object exposing ``node_requests``, ``node_responses``,
``mpc_lock`` and ``send_request`` whose class is named ``User`` to trigger auto-distribution (ml/module.py:326-346).
The reference's ``User`` spawns a network process and negotiates a job with a validator over TCP (nodes.py:380-414,
nodes/user_thread.py:242-392) — control plane, out of scope on one NVSwitch box where ranks come from ``torchrun``.
These shims keep the constructor contract so user scripts run unchanged.
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass


@dataclass
class UserConfig:
    """Field-compatible subset of the reference's ``UserConfig`` (nodes.py:16-77)."""
    upnp: bool = False
    off_chain_test: bool = True
    local_test: bool = True
    print_level: int = 0


class BaseNode:
    def __init__(self, config=None, **_):
        self.config = config
        self.node_requests: "queue.Queue" = queue.Queue()
        self.node_responses: "queue.Queue" = queue.Queue()
        self.mpc_lock = threading.Lock()

    def send_request(self, request_type, args=None, timeout=5):
        """nodes.py:201-235.  There is no network process: requests that only make sense across the WAN are
        answered locally; anything else is an error (the reference would silently time out)."""
        if request_type in ("debug_print", "release_memory"):
            return None
        raise NotImplementedError(f"local node shim cannot serve request {request_type!r}")

    def cleanup(self):
        pass


class User(BaseNode):
    pass


class Worker(BaseNode):
    pass
