"""HF-layout checkpoint reading (tensorlink_b200/ml/checkpoint.py) against directories written by ``transformers``
itself: config mapping, lazy per-tensor reads, sharded index.  CPU only (no kernels involved)."""
import json
import os

import torch
from safetensors.torch import save_file

from tensorlink_b200.ml import configs as C
from tensorlink_b200.ml.checkpoint import INDEX, LazyCheckpoint, config_from_dir, config_to_json
from tensorlink_b200.ml.weights import init_state_dict
from tests.hf_util import hf_model


def _same_cfg(a, b):
    keys = ("hidden", "intermediate", "n_layers", "n_heads", "n_kv_heads", "head_dim", "vocab", "tied", "qkv_bias", "qk_norm",
            "rope_theta", "rms_eps")
    return all(getattr(a, k) == getattr(b, k) for k in keys)


def test_reads_a_directory_written_by_transformers(tmp_path):
    for cfg in (C.TINY_QWEN2, C.TINY_QWEN3):
        d = tmp_path / cfg.name
        sd = init_state_dict(cfg, dtype=torch.bfloat16)
        hf_model(cfg, sd, "eager", torch.bfloat16).save_pretrained(d)
        assert _same_cfg(config_from_dir(str(d)), cfg)
        ck = LazyCheckpoint(str(d))
        assert ck.bytes_read == 0
        name = "model.layers.1.mlp.down_proj.weight"
        assert name in ck and "model.layers.99.mlp.down_proj.weight" not in ck
        assert torch.equal(ck[name], sd[name]) and ck.bytes_read == sd[name].numel() * 2      # nothing else was read
        for k in ck.keys():
            assert torch.equal(ck[k], sd[k]), k


def test_sharded_index_and_config_roundtrip(tmp_path):
    cfg = C.TINY_QWEN2_D128
    sd = init_state_dict(cfg, dtype=torch.bfloat16)
    names = sorted(sd)
    half = len(names) // 2
    files = {"model-00001-of-00002.safetensors": names[:half], "model-00002-of-00002.safetensors": names[half:]}
    for fn, ks in files.items():
        save_file({k: sd[k].contiguous() for k in ks}, str(tmp_path / fn))
    with open(tmp_path / INDEX, "w") as f:
        json.dump({"weight_map": {k: fn for fn, ks in files.items() for k in ks}}, f)
    with open(tmp_path / "config.json", "w") as f:
        json.dump(config_to_json(cfg), f)
    assert _same_cfg(config_from_dir(str(tmp_path)), cfg)
    ck = LazyCheckpoint(str(tmp_path))
    assert all(torch.equal(ck[k], sd[k]) for k in names)
    os.remove(tmp_path / INDEX)                       # without an index the files are scanned
    ck2 = LazyCheckpoint(str(tmp_path))
    assert set(ck2.keys()) == set(names) and torch.equal(ck2[names[-1]], sd[names[-1]])
