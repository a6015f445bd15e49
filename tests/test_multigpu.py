"""N = 2 GPUs: pipeline-sharded DistributedModel over NCCL equals the single-stage run (skipped with < 2 GPUs)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_pipeline_equals_single_stage(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
    errs = "".join(open(p).read() for p in sorted(map(str, tmp_path.glob("err*.txt"))))
    assert r.returncode == 0, errs or r.stderr[-4000:]
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert r0["logits_equal"] and r0["gen_equal"]
    assert r0["gen_graph_vs_eager"] and r1["gen_graph_vs_eager"]
    assert r0["used_ring"] and r1["used_ring"] and r0["gen_peer_vs_nccl"] and r1["gen_peer_vs_nccl"] and r0["stream_ok"]
    assert r0["bytes_sent_infer"] > 0 and r1["bytes_sent_infer"] > 0
    assert r0["eos_peer_equal"] and r0["eos_nccl_equal"]
    assert r0["eos_peer_shape"] == r1["eos_peer_shape"] and r0["eos_nccl_shape"] == r1["eos_nccl_shape"]
    assert abs(r0["loss"] - r0["loss_single"]) < 1e-5 and abs(r1["loss"] - r0["loss_single"]) < 1e-5
    ref = torch.load(tmp_path / "ref_grads.pt")
    for rank in (0, 1):
        g = torch.load(tmp_path / f"grads{rank}.pt")
        assert len(g) > 10
        for k, v in g.items():
            assert torch.equal(v, ref[k]), f"rank {rank} grad {k} differs from the single-stage run"
    t0, t1, tr = torch.load(tmp_path / "tied0.pt"), torch.load(tmp_path / "tied1.pt"), torch.load(tmp_path / "tied_ref.pt")
    assert torch.equal(t0, t1)                                        # both copies of the tied weight see the same gradient
    assert (t0.float() - tr.float()).norm() / tr.float().norm() < 5e-3   # = single-stage sum up to bf16 add order
    a0, a1, ar = torch.load(tmp_path / "tied2_0.pt"), torch.load(tmp_path / "tied2_1.pt"), torch.load(tmp_path / "tied2_ref.pt")
    assert torch.equal(a0, a1)                                        # still equal after an accumulated second backward
    assert (a0.float() - ar.float()).norm() / ar.float().norm() < 5e-3   # = twice the gradient, not 2(E+H)+E+H
    b0, b1 = torch.load(tmp_path / "tied3_0.pt"), torch.load(tmp_path / "tied3_1.pt")
    assert r1["tied_split_head"] and torch.equal(b0, b1)              # two micro-batches: split head backward, same exchange
    assert (b0.float() - tr.float()).norm() / tr.float().norm() < 8e-3


@pytest.mark.parametrize("world,rows", [(2, 2), (2, 5), (4, 2)])
def test_peer_ring_generation(tmp_path, world, rows):
    """Decode hops over peer-mapped mailboxes (first, middle and last stages) = NCCL hops = one stage, bit for bit;
    rows = 2 takes the GEMV path (the last GEMV stores into the neighbour), rows = 5 the GEMM path (the split-K reduce
    pass does)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "ring_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=dict(os.environ, PYTHONPATH=ROOT, RING_ROWS=str(rows)), capture_output=True, text=True, timeout=600)
    errs = "".join(open(p).read() for p in sorted(map(str, tmp_path.glob("err*.txt"))))
    assert r.returncode == 0, errs or r.stderr[-4000:]
    for rank in range(world):
        res = torch.load(tmp_path / f"ring{rank}.pt")
        assert res["used_ring"] and res["repeatable"] and res["peer_vs_nccl"], (rank, res)
    r0 = torch.load(tmp_path / "ring0.pt")
    assert r0["vs_single"] and r0["ckpt_roundtrip"] and len(r0["ckpt_files"]) == world
