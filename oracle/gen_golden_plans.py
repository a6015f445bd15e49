"""Golden shard plans produced by the reference's OWN planner (tensorlink/ml/graphing.py ``ModelParser``).  TEST INFRA.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden_plans
Writes tests/golden/ref_plans.json: for a few (model, workers, mode) cases the exact ``create_distributed_config`` result
(graphing.py:238-451) on a meta-device HF skeleton built from the hard-coded model constants (the reference builds the
same skeleton with ``load_model_skeleton``, utils.py:890-916, which needs the HF hub).  Cases are chosen so that the
decoder stack never fits one worker: the reference's loop finder does not match transformers-5.x layer loops,
which only matters when ``model.model`` as a whole could be assigned.
"""
import json
import os
import sys

import torch

from oracle.ref_shim import REFERENCE_ROOT, import_reference
from tensorlink_b200.ml import configs as C

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_plans.json")

CASES = [
    ("qwen25_7b_infer_8x3.2GB", "Qwen/Qwen2.5-7B", {f"w{i}": 3.2e9 for i in range(8)}, dict(training=False, max_seq_len=4096, batch_size=1)),
    ("qwen25_7b_infer_12GB_9GB", "Qwen/Qwen2.5-7B", {"a": 12e9, "b": 9e9}, dict(training=False, max_seq_len=2048, batch_size=1)),
    ("qwen3_8b_train_8x20GB", "Qwen/Qwen3-8B", {f"w{i}": 20e9 for i in range(8)}, dict(training=True, max_seq_len=1024, batch_size=1)),
    ("qwen25_05b_infer_0.8GB_0.9GB", "Qwen/Qwen2.5-0.5B", {"a": 0.8e9, "b": 0.9e9}, dict(training=False, max_seq_len=512, batch_size=1)),
    ("qwen25_7b_infer_hashed_workers", "Qwen/Qwen2.5-7B",
     {"509d89bf56704c67873c328e4f706a705b2fdc1671ebacab1083c9c6d2df650f": 10e9,
      "209d89bf56704c67873c328e4f706a705b2fdc1671ebacab1083c9c6d2df650f": 10e9}, dict(training=False, max_seq_len=1024, batch_size=4)),
]


def skeleton(cfg):
    if cfg.qk_norm:
        from transformers import Qwen3Config as HC, Qwen3ForCausalLM as M
        extra = dict(attention_bias=False)
    else:
        from transformers import Qwen2Config as HC, Qwen2ForCausalLM as M
        extra = {}
    hc = HC(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.intermediate, num_hidden_layers=cfg.n_layers,
            num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_kv_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_eps,
            max_position_embeddings=cfg.max_pos, tie_word_embeddings=cfg.tied, use_sliding_window=False, **extra)
    with torch.device("meta"):
        return M(hc).to(torch.bfloat16)


def main():
    import_reference()
    sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        import tensorlink.ml.graphing as G
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
    out = {}
    for tag, name, workers, kw in CASES:
        cfg = C.get_config(name)
        res = G.ModelParser(verbose=False).create_distributed_config(skeleton(cfg), {w: {"gpu_memory": m} for w, m in workers.items()},
                                                                     trusted=False, **kw)
        assert res["success"], tag
        out[tag] = {"model": name, "workers": workers, "kwargs": kw, "model_memory": res["model_memory"],
                    "host_memory_used": res["host_memory_used"], "config": res["config"]}
        print(tag, list(res["config"]))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=False)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
