"""Hard-coded model constants for the shard executor (no config.json is reachable offline).

The reference resolves these through ``AutoConfig.from_pretrained`` on the HF hub
(/root/reference/tensorlink/ml/utils.py:890-916 ``load_model_skeleton``); here they are the
public model-card values from SURVEY.md §8 and the parameter counts are asserted in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass, replace


@dataclass(frozen=True)
class ShardModelConfig:
    name: str
    hidden: int          # H
    intermediate: int    # I
    n_layers: int        # L
    n_heads: int         # n_h
    n_kv_heads: int      # n_kv
    head_dim: int        # d
    vocab: int           # V
    tied: bool           # lm_head shares embed_tokens
    qkv_bias: bool       # Qwen2: yes, Qwen3: no
    qk_norm: bool        # Qwen3: per-head RMSNorm on q/k before RoPE
    rope_theta: float = 1.0e6
    rms_eps: float = 1.0e-6
    max_pos: int = 32768

    @property
    def q_dim(self) -> int:
        return self.n_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.n_kv_heads * self.head_dim

    @property
    def qkv_dim(self) -> int:
        return self.q_dim + 2 * self.kv_dim

    def layer_matmul_params(self) -> int:
        """P_mm of SURVEY.md §8(d): matmul weights of one decoder layer."""
        return (self.hidden * self.qkv_dim + self.q_dim * self.hidden
                + 3 * self.hidden * self.intermediate)

    def layer_params(self) -> int:
        p = self.layer_matmul_params() + 2 * self.hidden
        if self.qkv_bias:
            p += self.qkv_dim
        if self.qk_norm:
            p += 2 * self.head_dim
        return p

    def total_params(self) -> int:
        p = self.n_layers * self.layer_params() + self.vocab * self.hidden + self.hidden
        if not self.tied:
            p += self.vocab * self.hidden
        return p

    def scaled(self, **kw) -> "ShardModelConfig":
        return replace(self, **kw)


QWEN25_05B = ShardModelConfig("Qwen/Qwen2.5-0.5B", 896, 4864, 24, 14, 2, 64, 151936,
                              tied=True, qkv_bias=True, qk_norm=False)
QWEN25_7B = ShardModelConfig("Qwen/Qwen2.5-7B", 3584, 18944, 28, 28, 4, 128, 152064,
                             tied=False, qkv_bias=True, qk_norm=False)
QWEN25_7B_INSTRUCT = replace(QWEN25_7B, name="Qwen/Qwen2.5-7B-Instruct")
QWEN3_8B = ShardModelConfig("Qwen/Qwen3-8B", 4096, 12288, 36, 32, 8, 128, 151936,
                            tied=False, qkv_bias=False, qk_norm=True, max_pos=40960)

# Small same-architecture configs used by parity tests (oracle finishes in seconds on CPU).
TINY_QWEN2 = ShardModelConfig("tiny-qwen2", 256, 768, 4, 4, 2, 64, 1024,
                              tied=True, qkv_bias=True, qk_norm=False, max_pos=4096)
TINY_QWEN2_D128 = ShardModelConfig("tiny-qwen2-d128", 512, 1536, 4, 4, 2, 128, 2048,
                                   tied=False, qkv_bias=True, qk_norm=False, max_pos=4096)
TINY_QWEN3 = ShardModelConfig("tiny-qwen3", 512, 1024, 4, 4, 2, 128, 2048,
                              tied=False, qkv_bias=False, qk_norm=True, max_pos=4096)

REGISTRY = {c.name: c for c in (QWEN25_05B, QWEN25_7B, QWEN25_7B_INSTRUCT, QWEN3_8B,
                                TINY_QWEN2, TINY_QWEN2_D128, TINY_QWEN3)}


def get_config(name: str) -> ShardModelConfig:
    try:
        return REGISTRY[name]
    except KeyError as e:
        raise KeyError(f"unknown model {name!r}; known: {sorted(REGISTRY)}") from e
