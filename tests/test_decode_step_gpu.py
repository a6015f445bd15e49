"""The persistent decode-step kernel (csrc/decode_step.cu) against the per-kernel launch sequence it replaces:
same arithmetic and accumulation order, so hidden states, logits, KV cache and ids must be IDENTICAL."""
import os

import pytest
import torch

from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.weights import synthetic_tokens

pytestmark = pytest.mark.gpu


def _run(cfg, B, prompt, new, impl, **kw):
    from tensorlink_b200.ml import DistributedModel
    from tensorlink_b200.ml import shard
    os.environ["TL_DECODE_IMPL"] = impl
    saved, shard._GEMV_MAX_ROWS = shard._GEMV_MAX_ROWS, 8      # the step kernel restates the GEMV launch sequence (rows <= 4)
    try:
        dm = DistributedModel(cfg, training=False, max_batch=B, max_seq=prompt + new + 3, **kw)
        ids = synthetic_tokens(cfg, B, prompt)
        out = dm.generate(ids, max_new_tokens=new).cpu()
        st = dm.stage
        return out, st.logits_dec[:B].cpu().clone(), [k.cpu().clone() for k in st.slots[0].kc], st.n_decode_launches(B)
    finally:
        os.environ.pop("TL_DECODE_IMPL", None)
        shard._GEMV_MAX_ROWS = saved


@pytest.mark.parametrize("cfg", [C.TINY_QWEN2, C.TINY_QWEN2_D128, C.TINY_QWEN3], ids=lambda c: c.name)
@pytest.mark.parametrize("B", [1, 2, 4])
def test_step_kernel_equals_kernel_sequence(cfg, B):
    a = _run(cfg, B, 11, 20, "kernels")
    b = _run(cfg, B, 11, 20, "step")
    assert b[3] == 1 and a[3] > 1
    assert torch.equal(a[0], b[0])                       # generated ids
    assert torch.equal(a[1], b[1])                       # logits of the last decode step, bit for bit
    for ka, kb in zip(a[2], b[2]):
        assert torch.equal(ka, kb)                       # KV cache


def test_step_kernel_full_size_05b():
    a = _run(C.QWEN25_05B, 1, 32, 16, "kernels", init="device")
    b = _run(C.QWEN25_05B, 1, 32, 16, "step", init="device")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_step_kernel_repeated_launches_leave_sync_clean():
    from tensorlink_b200.ml import DistributedModel
    cfg = C.TINY_QWEN2_D128
    os.environ["TL_DECODE_IMPL"] = "step"
    try:
        dm = DistributedModel(cfg, training=False, max_batch=2, max_seq=64)
        ids = synthetic_tokens(cfg, 2, 9)
        x = dm.generate(ids, max_new_tokens=30).cpu()
        y = dm.generate(ids, max_new_tokens=30).cpu()
        z = dm.generate(ids, max_new_tokens=30, use_graph=False).cpu()
    finally:
        os.environ.pop("TL_DECODE_IMPL", None)
    assert torch.equal(x, y) and torch.equal(x, z)
    assert int(dm.stage.step_ws[:8].view(torch.int32).abs().sum()) == 0
