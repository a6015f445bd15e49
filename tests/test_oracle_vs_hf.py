"""Pin the oracle against the installed Hugging Face implementation on CPU."""
import pytest
import torch

from oracle import shard_oracle as O
from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.weights import init_state_dict, synthetic_tokens
from tests.hf_util import hf_model

CASES = [C.TINY_QWEN2, C.TINY_QWEN2_D128, C.TINY_QWEN3]


def test_param_counts():
    assert C.QWEN25_05B.layer_params() == 14_912_384
    assert C.QWEN25_05B.total_params() == 494_032_768
    assert C.QWEN25_7B.layer_params() == 233_057_792
    assert C.QWEN25_7B.layer_matmul_params() == 233_046_016
    assert C.QWEN25_7B.total_params() == 7_615_616_512
    assert C.QWEN3_8B.layer_params() == 192_946_432
    assert C.QWEN3_8B.total_params() == 8_190_735_360


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_logits_bit_exact_vs_hf_eager(cfg, dtype):
    sd = init_state_dict(cfg, dtype=dtype)
    ids = synthetic_tokens(cfg, 2, 24)
    hf = hf_model(cfg, sd, "eager", dtype)
    with torch.no_grad():
        ref = hf(input_ids=ids).logits
        got = O.OracleModel(cfg, sd, "eager").logits(ids)
        got2 = O.OracleModel(cfg, sd, "eager").logits(ids, n_shards=3)
    assert torch.equal(got, ref)
    assert torch.equal(got2, ref)     # sharded == unsharded (injector.py:154-281 contract)


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: c.name)
def test_sdpa_math_close_to_hf_sdpa(cfg):
    """'sdpa_math' restates the SDPA contract, not torch's CPU flash kernel blocking, so it is
    compared with HF(sdpa) relative to the spread between HF's own two attention paths
    (eager vs sdpa), i.e. the bf16 noise floor of the reference itself."""
    sd = init_state_dict(cfg)
    ids = synthetic_tokens(cfg, 2, 40)
    with torch.no_grad():
        ref = hf_model(cfg, sd, "sdpa")(input_ids=ids).logits
        ref_eager = hf_model(cfg, sd, "eager")(input_ids=ids).logits
        got = O.OracleModel(cfg, sd, "sdpa_math").logits(ids)
    floor = O.rel_l2(ref_eager, ref)
    assert O.rel_l2(got, ref) <= 1.25 * floor, (O.rel_l2(got, ref), floor)


def test_attention_op_sdpa_math_vs_torch_sdpa():
    """Per-op: against fp32 SDPA on the same bf16 inputs the restatement is as accurate as
    torch's own bf16 CPU kernel (both ~2e-3: one bf16 rounding of P and of the output)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    B, nh, nkv, S, d = 2, 4, 2, 40, 64
    q = torch.randn(B, nh, S, d, generator=g).bfloat16()
    k = torch.randn(B, nkv, S, d, generator=g).bfloat16()
    v = torch.randn(B, nkv, S, d, generator=g).bfloat16()
    kk, vv = O.repeat_kv(k, 2), O.repeat_kv(v, 2)
    f32 = F.scaled_dot_product_attention(q.float(), kk.float(), vv.float(), is_causal=True)
    f32 = f32.transpose(1, 2).reshape(B, S, -1)
    tb = F.scaled_dot_product_attention(q, kk, vv, is_causal=True).transpose(1, 2).reshape(B, S, -1)
    mine = O.attention_sdpa_math(q, k, v, d ** -0.5, 2)
    assert O.rel_l2(mine, f32) <= 1.25 * O.rel_l2(tb, f32)


@pytest.mark.parametrize("cfg", [C.TINY_QWEN2, C.TINY_QWEN3], ids=lambda c: c.name)
def test_greedy_generate_ids_match_hf(cfg):
    sd = init_state_dict(cfg, dtype=torch.float32)
    ids = synthetic_tokens(cfg, 2, 8)
    hf = hf_model(cfg, sd, "eager", torch.float32)
    with torch.no_grad():
        ref = hf.generate(ids, max_new_tokens=12, do_sample=False, eos_token_id=None, pad_token_id=0)
    got = O.OracleModel(cfg, sd, "eager").generate(ids, 12, n_shards=2)
    assert torch.equal(got, ref)


def test_loss_and_grads_match_hf():
    cfg = C.TINY_QWEN2
    sd = {k: v.clone().requires_grad_(True) for k, v in init_state_dict(cfg, dtype=torch.float32).items()
          if k != "lm_head.weight"}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    ids = synthetic_tokens(cfg, 2, 16)
    hf = hf_model(cfg, {k: v.detach() for k, v in sd.items()}, "eager", torch.float32).train()
    out = hf(input_ids=ids, labels=ids)
    out.loss.backward()
    loss, _ = O.OracleModel(cfg, sd, "eager").loss(ids, ids, n_shards=1)
    loss.backward()
    assert torch.allclose(loss, out.loss, rtol=0, atol=1e-6)
    g_ref = dict(hf.named_parameters())
    for n in ("model.layers.0.self_attn.q_proj.weight", "model.layers.3.mlp.down_proj.weight",
              "model.embed_tokens.weight", "model.layers.1.input_layernorm.weight"):
        assert torch.allclose(sd[n].grad, g_ref[n].grad, rtol=1e-4, atol=1e-6), n
