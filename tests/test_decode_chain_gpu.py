"""The persistent decode-chain kernel (csrc/decode_chain.cu) against the per-kernel launch sequence and the oracle.

The GEMV jobs of a chain are bit-identical to ``gemv_stream_kernel``; the attention job is split-KV over all CTAs
(another summation order than the one-CTA-per-head kernel), so whole steps are compared (a) with the per-kernel path
within the attention tolerance of tests/test_kernels_gpu.py, (b) with the CPU oracle under the chain criteria, and
(c) on greedy ids wherever the oracle's margin is resolvable."""
import pytest
import torch

from oracle import shard_oracle as O
from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.weights import init_state_dict, synthetic_tokens

pytestmark = pytest.mark.gpu
MARGIN = 0.05


def _dm(cfg, monkeypatch, impl, layers_per_launch=1, **kw):
    from tensorlink_b200.ml import DistributedModel
    monkeypatch.setenv("TL_DECODE_IMPL", impl)
    monkeypatch.setenv("TL_CHAIN_LAYERS", str(layers_per_launch))
    kw.setdefault("max_seq", 128)
    return DistributedModel(cfg, training=False, **kw)


def _decode_logits(dm, ids, steps):
    """prefill + `steps` teacher-forced decode steps; returns the bf16 logits of every decode step [steps, B, V]."""
    st = dm.stage
    B, S = ids.shape
    x = st.prefill(st.embed(ids[:, :S - steps].cuda()), 0, 0)
    out = []
    for s in range(steps):
        st.ids_dec[0][:B].copy_(ids[:, S - steps + s].cuda())
        st.decode(0, B, use_graph=(s % 2 == 1))
        out.append(st.logits_dec[:B].clone())
    st.check()
    return torch.stack(out).cpu()


@pytest.mark.parametrize("cfg", [C.TINY_QWEN2, C.TINY_QWEN2_D128, C.TINY_QWEN3], ids=lambda c: c.name)
@pytest.mark.parametrize("B", [1, 2, 3])
def test_chain_step_equals_kernel_sequence_and_oracle(cfg, B, monkeypatch):
    ids = synthetic_tokens(cfg, B, 40)
    steps = 6
    chain = _decode_logits(_dm(cfg, monkeypatch, "chain", max_batch=B), ids, steps)
    kern = _decode_logits(_dm(cfg, monkeypatch, "kernels", max_batch=B), ids, steps)
    sd = init_state_dict(cfg)
    # decode step s consumed token S-steps+s, i.e. it predicts position S-steps+s+1: logits row S-steps+s of the one-shot forward
    with torch.no_grad():
        full16 = O.OracleModel(cfg, sd, "sdpa_math").logits(ids)
        full32 = O.OracleModel(cfg, {k: v.float() for k, v in sd.items()}, "sdpa_math").logits(ids)
    S = ids.shape[1]
    ref16 = full16[:, S - steps:S].transpose(0, 1)
    ref32 = full32[:, S - steps:S].transpose(0, 1)
    e_ref = O.rel_l2(ref16, ref32)
    e_chain, e_kern, mutual = O.rel_l2(chain, ref32), O.rel_l2(kern, ref32), O.rel_l2(chain, kern)
    print(f"{cfg.name} B={B}: chain-vs-fp32 {e_chain:.3e} kernels-vs-fp32 {e_kern:.3e} oracle_bf16-vs-fp32 {e_ref:.3e} chain-vs-kernels {mutual:.3e}")
    assert e_chain <= 1.25 * e_ref and O.rel_l2(chain, ref16) <= 2.0 * e_ref
    assert mutual <= 2.0 * e_ref
    top2 = ref32.float().topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > MARGIN
    assert torch.equal(chain.float().argmax(-1)[safe], ref32.argmax(-1)[safe])


@pytest.mark.parametrize("layers_per_launch", [2, 3])
def test_multi_layer_chains_are_bit_identical_to_single_layer_chains(layers_per_launch, monkeypatch):
    """Grouping more layers into one launch changes no arithmetic (same jobs, same order, same partitioning)."""
    cfg = C.TINY_QWEN2_D128
    ids = synthetic_tokens(cfg, 2, 24)
    a = _dm(cfg, monkeypatch, "chain", 1, max_batch=2).generate(ids, max_new_tokens=20)
    b = _dm(cfg, monkeypatch, "chain", layers_per_launch, max_batch=2).generate(ids, max_new_tokens=20)
    assert torch.equal(a, b)


def test_chain_graph_replay_equals_eager_and_is_repeatable(monkeypatch):
    """The sync slots clean themselves: a second generation (eager, then graph replay) reproduces the first."""
    cfg = C.TINY_QWEN3
    ids = synthetic_tokens(cfg, 1, 16)
    dm = _dm(cfg, monkeypatch, "chain", max_batch=1)
    a = dm.generate(ids, max_new_tokens=32)
    b = dm.generate(ids, max_new_tokens=32, use_graph=False)
    c = dm.generate(ids, max_new_tokens=32)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert int(dm.stage.slots[0].chain_sync[:, :3].abs().sum()) == 0        # barrier / exit / error words back to zero


@pytest.mark.parametrize("max_seq,S", [(2048, 1990), (4096, 4040)])
def test_chain_long_context_full_size(max_seq, S, monkeypatch):
    """Full-size Qwen2.5-0.5B at BASELINE config 3 / 5 context lengths: the chain's split-KV attention over thousands of
    cached keys picks the argmax of ONE forward over the final sequence wherever that forward's margin is resolvable,
    and agrees with the per-kernel path's logits."""
    cfg = C.QWEN25_05B
    ids = synthetic_tokens(cfg, 1, S)
    dm = _dm(cfg, monkeypatch, "chain", max_batch=1, max_seq=max_seq, init="device")
    out = dm.generate(ids, max_new_tokens=16).cpu()
    logits = dm(out[:, :-1]).logits[:, S - 1:].cpu().float()
    top2 = logits.topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > MARGIN
    assert safe.float().mean() > 0.2
    assert torch.equal(logits.argmax(-1)[safe], out[:, S:][safe])
    dk = _dm(cfg, monkeypatch, "kernels", max_batch=1, max_seq=max_seq, init="device")
    out_k = dk.generate(ids, max_new_tokens=16).cpu()
    first_unsafe = int((~safe[0]).nonzero()[0]) if (~safe[0]).any() else 16
    assert torch.equal(out[0, S:S + first_unsafe], out_k[0, S:S + first_unsafe])
