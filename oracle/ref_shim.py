"""Import the UNMODIFIED reference (``/root/reference/tensorlink``) in this container.  TEST INFRA.

The reference needs ``accelerate``, ``web3``, ``miniupnpc``, ``eth_abi``, ``hexbytes`` (absent here) and a
transformers 4.x symbol; none of them is touched by the hot path, so they are stubbed.
The reference creates ``logs/``, ``tmp/`` and ``keys/`` in the CWD at import time, so the import runs from a
scratch directory.  Only ``oracle/gen_golden.py`` uses this, and only where /root/reference exists.
"""
import contextlib
import importlib.machinery
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present (expected only in the build container)")
    import torch
    import transformers  # noqa: F401  (must be imported before the accelerate stub exists)

    @contextlib.contextmanager
    def init_empty_weights(include_buffers=False):
        with torch.device("meta"):
            yield

    _stub("accelerate", init_empty_weights=init_empty_weights)
    _stub("miniupnpc", UPnP=type("UPnP", (), {}))
    w3 = _stub("web3", Web3=type("Web3", (), {"keccak": staticmethod(lambda *a, **k: b""),
                                               "HTTPProvider": staticmethod(lambda *a, **k: None)}))
    w3.__path__ = []
    _stub("web3.exceptions", ContractLogicError=type("ContractLogicError", (Exception,), {}))
    _stub("eth_abi", encode=lambda *a, **k: b"")
    _stub("hexbytes", HexBytes=bytes)
    if not hasattr(transformers, "AutoModelForVision2Seq"):
        transformers.AutoModelForVision2Seq = transformers.AutoModelForImageTextToText
    scratch = tempfile.mkdtemp(prefix="tlref_")
    cwd = os.getcwd()
    os.chdir(scratch)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import tensorlink.ml.injector as injector
        import tensorlink.ml.utils as utils
    finally:
        sys.path.remove(REFERENCE_ROOT)
        os.chdir(cwd)
    return injector, utils
