"""bench.py's reference arm (the CPU leg the driver runs as `--impl reference`) prints ONE JSON line with the contract's
keys.  Runs the tiny workload so that it takes seconds; no GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                        "--warmup", "1", "--train-model", "tiny-qwen2-d128", "--train-seq", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    hop = cb["reference_hop"]
    assert hop["decode"]["seconds"] > 0 and hop["prefill"]["payload_bytes"] > hop["decode"]["payload_bytes"]
    # the training half of BASELINE's metric travels on the same line
    tr = d["train"]
    assert tr["unit"] == "samples/s" and tr["value"] > 0 and tr["cpu_baseline"]["kind"] == "port" and tr["cpu_baseline"]["cores"] >= 1
