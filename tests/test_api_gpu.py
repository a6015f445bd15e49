"""API-surface behaviour of the drop-in front end: HF module in, state dict out, edge cases, loud errors."""
import pytest
import torch

from oracle import shard_oracle as O
from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.weights import init_state_dict, synthetic_tokens
from tests.hf_util import hf_model

pytestmark = pytest.mark.gpu


def test_hf_module_in_state_dict_out():
    """DistributedModel(model=<HF nn.Module>) (module.py:251-265 accepts a module): weights are taken from the module;
    state_dict() returns them under HF names, bit for bit, and an HF model loaded from it reproduces the logits."""
    from tensorlink_b200.ml import DistributedModel
    for cfg in (C.TINY_QWEN2, C.TINY_QWEN3):
        sd = init_state_dict(cfg)
        hf = hf_model(cfg, sd, "sdpa")
        dm = DistributedModel(hf, training=False, max_batch=2, max_seq=64)
        assert dm.cfg.hidden == cfg.hidden and dm.cfg.qk_norm == cfg.qk_norm and dm.cfg.tied == cfg.tied
        out_sd = dm.state_dict()
        for k, v in sd.items():
            assert torch.equal(out_sd[k].cpu(), v), k
        ids = synthetic_tokens(cfg, 2, 20)
        with torch.no_grad():
            ref = hf(input_ids=ids).logits
            ref32 = O.OracleModel(cfg, {k: v.float() for k, v in sd.items()}, "sdpa_math").logits(ids)
        got = dm(input_ids=ids).logits.cpu()
        e_ref, e_gpu = O.rel_l2(ref, ref32), O.rel_l2(got, ref32)
        assert e_gpu <= 1.25 * e_ref, (e_gpu, e_ref)
        assert len(list(dm.parameters())) == len(out_sd)


def test_generate_edge_cases():
    from tensorlink_b200.ml import DistributedModel
    cfg = C.TINY_QWEN2_D128
    dm = DistributedModel(cfg, training=False, max_batch=3, max_seq=48)
    one = synthetic_tokens(cfg, 1, 1)
    g = dm.generate(one, max_new_tokens=1)
    assert g.shape == (1, 2) and int(g[0, 0]) == int(one[0, 0])
    g3 = dm.generate(synthetic_tokens(cfg, 3, 7), max_new_tokens=5)          # odd batch (GEMV path with M = 3)
    assert g3.shape == (3, 12)
    full = dm.generate(synthetic_tokens(cfg, 1, 40), max_new_tokens=8)       # fills the cache exactly (40 + 8 = 48)
    assert full.shape == (1, 48)
    with pytest.raises(ValueError):
        dm.generate(synthetic_tokens(cfg, 1, 41), max_new_tokens=8)          # would overflow the KV cache: loud
    with pytest.raises(ValueError):
        dm.generate(synthetic_tokens(cfg, 4, 4), max_new_tokens=2)           # more rows than the stage was sized for
    with pytest.raises(NotImplementedError):
        dm.generate(one, max_new_tokens=2, num_beams=4)                      # unsupported HF keywords are refused, not dropped
    with pytest.raises(NotImplementedError):
        dm(one, position_ids=torch.zeros(1, 1, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        dm(synthetic_tokens(cfg, 2, 4), attention_mask=torch.tensor([[0, 1, 1, 1], [1, 1, 1, 1]]))   # forward: no padded rows
    assert dm.generate(one, max_new_tokens=2, use_cache=True, num_beams=1).shape == (1, 3)           # neutral values pass
    with pytest.raises(NotImplementedError):
        DistributedModel(cfg, training=False, dtype=torch.float32)


def test_generate_left_padded_batch_and_eos_early_stop():
    """Batched-generation conventions of HF: a left-padded batch with its attention_mask (every row attends to its own
    tokens at positions 0..L-1; the result keeps the pads in front) and stopping once every row has emitted EOS."""
    from tensorlink_b200.ml import DistributedModel
    cfg = C.TINY_QWEN2_D128
    sd = init_state_dict(cfg)
    dm = DistributedModel(cfg, training=False, max_batch=4, max_seq=96)
    lens, S, new, PAD = [9, 14, 9, 5], 14, 12, 7
    rows = [synthetic_tokens(cfg, 1, L, seed=100 + i)[0] for i, L in enumerate(lens)]
    ids = torch.full((4, S), PAD, dtype=torch.int64)
    mask = torch.zeros(4, S, dtype=torch.int64)
    for r, (t, L) in enumerate(zip(rows, lens)):
        ids[r, S - L:], mask[r, S - L:] = t, 1
    out = dm.generate(ids, attention_mask=mask, max_new_tokens=new, pad_token_id=PAD).cpu()
    assert out.shape == (4, S + new) and torch.equal(out[:, :S], ids)
    oracle = O.OracleModel(cfg, sd, "sdpa_math")
    for r, (t, L) in enumerate(zip(rows, lens)):
        alone = dm.generate(t[None], max_new_tokens=new).cpu()                # the same row without padding
        want, margins = oracle.generate(t[None], new, return_margins=True)
        for s in range(new):                                                  # exact until the first unresolvable step
            if margins[0, s] < 0.05:
                break
            assert out[r, S + s] == want[0, L + s] == alone[0, L + s], (r, s)
    with pytest.raises(NotImplementedError):
        dm.generate(ids, attention_mask=mask.flip(1), max_new_tokens=2)       # right padding is not a generation layout
    # ---- EOS: take the token the model emits at step 2 as the EOS id; HF semantics = the result ends right after it
    base = dm.generate(rows[1][None], max_new_tokens=40).cpu()
    eos = int(base[0, 14 + 2])
    got = dm.generate(rows[1][None], max_new_tokens=40, eos_token_id=eos).cpu()
    first = int((base[0, 14:] == eos).nonzero()[0])
    assert torch.equal(got, base[:, :14 + first + 1])
    # the decode loop stopped at the first check after the EOS instead of running all 40 steps
    assert int(dm.stage.slots[0].pos_dev.item()) <= 14 + 16


def test_forward_kwargs_and_train_eval_switch():
    from tensorlink_b200.ml import DistributedModel
    cfg = C.TINY_QWEN2
    dm = DistributedModel(cfg, training=True, max_batch=2, max_seq=32, optimizer=torch.optim.AdamW)
    ids = synthetic_tokens(cfg, 2, 16)
    dm.eval()
    a = dm(ids).logits                       # positional (module.py:355-359)
    b = dm(input_ids=ids).logits             # keyword
    assert torch.equal(a, b)
    dm.train()
    with pytest.raises(ValueError):
        dm(ids)                              # training forward needs labels
    out = dm(ids, labels=ids)
    assert out.logits is None and out.loss.requires_grad
    out.loss.backward()
    opt = dm.create_optimizer(lr=1e-3, weight_decay=0.1)
    opt.step()
    opt.zero_grad()
    # gradients read through the export are all zero (matrix gradients are zeroed lazily: the first weight-gradient GEMM
    # of the next step writes them, and any export in between settles them first)
    assert all(float(t.float().abs().sum()) == 0.0 for t in dm.stage.params.hf_state_dict(grads=True).values())
    out2 = dm(ids, labels=ids)
    out2.loss.backward()                     # a second step accumulates into clean gradients
    g2 = dm.stage.params.hf_state_dict(grads=True)
    assert any(float(t.float().abs().sum()) > 0 for t in g2.values())


@pytest.mark.parametrize("cfg", [C.TINY_QWEN2, C.TINY_QWEN2_D128], ids=lambda c: c.name)
def test_checkpoint_roundtrip_through_hf_layout(tmp_path, cfg):
    """save_pretrained -> a directory `transformers` itself can load, and DistributedModel(<dir>) reads it back lazily:
    same logits bit for bit, only this stage's tensors touched (tied head stored once)."""
    from transformers import AutoModelForCausalLM
    from tensorlink_b200.ml import DistributedModel
    from tensorlink_b200.ml.checkpoint import LazyCheckpoint
    ids = synthetic_tokens(cfg, 2, 12)
    a = DistributedModel(cfg, training=False, max_batch=2, max_seq=64, seed=99)
    want = a(ids).logits.cpu()
    d = str(tmp_path / "ckpt")
    a.save_pretrained(d)
    hf = AutoModelForCausalLM.from_pretrained(d, dtype=torch.bfloat16)
    ref_sd = a.state_dict()
    for k, v in hf.state_dict().items():
        assert torch.equal(v, ref_sd[k].cpu()), k
    b = DistributedModel(d, training=False, max_batch=2, max_seq=64, seed=1)          # seed differs: weights come from disk
    assert torch.equal(b(ids).logits.cpu(), want)
    ck = LazyCheckpoint(d)
    assert ("lm_head.weight" in ck) == (not cfg.tied)
