"""Build installed-HF models from a ShardModelConfig (test helper; pins the oracle)."""
import torch


def hf_model(cfg, sd, attn="eager", dtype=torch.bfloat16):
    if cfg.qk_norm:
        from transformers import Qwen3Config as C, Qwen3ForCausalLM as M
        extra = dict(attention_bias=False)
    else:
        from transformers import Qwen2Config as C, Qwen2ForCausalLM as M
        extra = {}
    hc = C(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.intermediate,
           num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_heads,
           num_key_value_heads=cfg.n_kv_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_eps,
           max_position_embeddings=cfg.max_pos, tie_word_embeddings=cfg.tied,
           rope_parameters={"rope_type": "default", "rope_theta": cfg.rope_theta},
           use_sliding_window=False, attn_implementation=attn, **extra)
    m = M(hc)
    inv = m.model.rotary_emb.inv_freq.clone()      # from_pretrained keeps this buffer fp32
    m = m.to(dtype)
    m.model.rotary_emb.inv_freq = inv
    m.model.rotary_emb.original_inv_freq = inv.clone()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    if cfg.tied:
        assert m.lm_head.weight.data_ptr() == m.model.embed_tokens.weight.data_ptr()
    return m.eval()
