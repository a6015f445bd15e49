"""GPU diagnostics: per-op error decomposition of one decoder layer against the oracle trace."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import shard_oracle as O
from tensorlink_b200 import native as nat
from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.shard import CudaLayerGroup, ShardParams
from tensorlink_b200.ml.weights import init_state_dict, synthetic_tokens


def attn_diag():
    for std in (1.0, 0.3, 0.1):
        for (B, S, past, n_h, n_kv, d) in [(2, 200, 0, 14, 2, 64), (1, 96, 160, 28, 4, 128)]:
            g = torch.Generator().manual_seed(60)
            T = past + S
            q = (torch.randn(B, S, n_h, d, generator=g) * std).bfloat16()
            k = (torch.randn(B, n_kv, T, d, generator=g) * std).bfloat16()
            v = torch.randn(B, n_kv, T, d, generator=g).bfloat16()
            ref = O.attention_sdpa_math(q.transpose(1, 2), k, v, d ** -0.5, n_h // n_kv)
            kk, vv = O.repeat_kv(k, n_h // n_kv).float(), O.repeat_kv(v, n_h // n_kv).float()
            s = (q.transpose(1, 2).float() @ kk.transpose(2, 3)) * d ** -0.5 + O.causal_mask(S, T, torch.float32)
            f32 = (F.softmax(s, -1) @ vv).transpose(1, 2).reshape(B, S, -1)
            out = torch.empty(B, S, n_h * d, dtype=torch.bfloat16, device="cuda")
            nat.attn_prefill_fwd(q.cuda(), k.cuda().contiguous(), v.cuda().contiguous(), out, None, B, S, past, n_h, n_kv, d, d ** -0.5)
            got = out.cpu()
            print(f"attn std={std} d={d} S={S} past={past}: gpu-f32 {O.rel_l2(got, f32):.2e} ref-f32 {O.rel_l2(ref, f32):.2e} "
                  f"gpu-ref {O.rel_l2(got, ref):.2e} round(f32)-f32 {O.rel_l2(f32.bfloat16(), f32):.2e} "
                  f"gpu-round(f32) {O.rel_l2(got, f32.bfloat16()):.2e} ref-round(f32) {O.rel_l2(ref, f32.bfloat16()):.2e}")


def layer_diag(cfg):
    sd = init_state_dict(cfg)
    ids = synthetic_tokens(cfg, 2, 33)
    m = O.OracleModel(cfg, sd, "sdpa_math")
    B, S = ids.shape
    with torch.no_grad():
        x0 = F.embedding(ids, m.embed)
        cos, sin = O.rope_tables(cfg, torch.arange(S)[None].expand(B, -1), x0.dtype)
        tr = {}
        y = O.decoder_layer(cfg, m.layers[0], x0, cos, sin, "sdpa_math", trace=tr)
        y32 = O.decoder_layer(cfg, O.LayerWeights(**{k: (v.float() if v is not None else None) for k, v in m.layers[0].__dict__.items()}),
                              x0.float(), cos.float(), sin.float(), "sdpa_math")
    p = ShardParams(cfg, [0], False, False, "cuda")
    p.load_hf_state_dict(sd)
    grp = CudaLayerGroup(cfg, p, B, 64)
    N = B * S
    w = grp._bufs(N)
    v = p.v
    x = x0.cuda().reshape(N, -1).contiguous()
    nat.rmsnorm_fwd(x, v["l0.ln1"], cfg.rms_eps, out=w.h)
    print(cfg.name, "ln1", O.rel_l2(w.h.cpu().view(B, S, -1), tr["ln1"]))
    nat.gemm(w.h, v["l0.wqkv"], out=w.qkv, bias=v.get("l0.bqkv"))
    h = tr["ln1"]
    lw = m.layers[0]
    qkv_ref = torch.cat([F.linear(h, lw.wq, lw.bq), F.linear(h, lw.wk, lw.bk), F.linear(h, lw.wv, lw.bv)], -1)
    print(cfg.name, "qkv", O.rel_l2(w.qkv.cpu().view(B, S, -1), qkv_ref))
    grp.pos_dev.fill_(0)
    nat.rope_kv_fwd(w.qkv, w.q, grp.kc[0], grp.vc[0], grp.pos_dev, grp.cos, grp.sin, v.get("l0.qn"), v.get("l0.kn"), cfg.rms_eps, S,
                    cfg.n_heads, cfg.n_kv_heads, cfg.head_dim)
    print(cfg.name, "q_rope", O.rel_l2(w.q.cpu().view(B, S, cfg.n_heads, -1).transpose(1, 2), tr["q_rope"]))
    print(cfg.name, "k_rope", O.rel_l2(grp.kc[0].cpu()[:, :, :S], tr["k_rope"]))
    nat.attn_prefill_fwd(w.q, grp.kc[0], grp.vc[0], w.attn, None, B, S, 0, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, grp.scale)
    print(cfg.name, "attn", O.rel_l2(w.attn.cpu().view(B, S, -1), tr["attn"]))
    nat.gemm(w.attn, v["l0.wo"], out=x, residual=x)
    print(cfg.name, "post_attn", O.rel_l2(x.cpu().view(B, S, -1), tr["post_attn"]))
    nat.rmsnorm_fwd(x, v["l0.ln2"], cfg.rms_eps, out=w.h)
    nat.gemm(w.h, v["l0.wgu"], out=w.act, flags=nat.EPI_SWIGLU)
    pa = tr["post_attn"]
    h2 = O.rmsnorm(pa, lw.ln2, cfg.rms_eps)
    act_ref = F.silu(F.linear(h2, lw.wg)) * F.linear(h2, lw.wu)
    print(cfg.name, "act", O.rel_l2(w.act.cpu().view(B, S, -1), act_ref))
    nat.gemm(w.act, v["l0.wd"], out=x, residual=x)
    got = x.cpu().view(B, S, -1)
    print(cfg.name, "layer out gpu-ref", O.rel_l2(got, y), "gpu-f32", O.rel_l2(got, y32), "ref-f32", O.rel_l2(y, y32),
          "|x0|", float(x0.float().norm()), "|y-x0|", float((y.float() - x0.float()).norm()))


if __name__ == "__main__":
    nat.require_device()
    attn_diag()
    for cfg in (C.TINY_QWEN2, C.TINY_QWEN3):
        layer_diag(cfg)
