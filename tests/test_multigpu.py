"""Multi-GPU checks on real hardware (skipped on a one-GPU box): the NCCL gather and the strong-scaling path, each as a
torchrun job over all visible GPUs.  The same logic runs over gloo with world_size 2 in the CPU suite
(tests/test_pipeline.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script, port, *args):
    n = min(_ngpu(), 8)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)] + list(args),
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nccl_gather_matches_single_gpu_decode():
    r = _torchrun("gpu_gather_check.py", 29551)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_strong_scaling_shards_match_single_gpu_decode_and_oracle_at_seams():
    r = _torchrun("gpu_strong_check.py", 29552)
    print(r.stdout[-400:])
    assert r.returncode == 0 and "STRONG_OK" in r.stdout, r.stdout[-3000:]
