"""GPU, two or more devices: the data-parallel step with the gradient all-reduce fused into the finalize kernel (peer memory over
NVLink) against a single-process run.  Skipped on single-GPU boxes; the host-side sharding logic is covered on CPU with gloo
(tests/test_distributed_cpu.py)."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("force_nccl", [False, True], ids=["peer-memory", "nccl"])
def test_data_parallel_step_matches_single_process(force_nccl):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env.pop("DRK_NO_PEER_EXCHANGE", None)
    if force_nccl:
        env["DRK_NO_PEER_EXCHANGE"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29541" if force_nccl else "29540", os.path.join(ROOT, "tests", "_peer_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    assert out.count("PEER_OK") == 2, out[-3000:]
    assert ("mode nccl" in out) == force_nccl or "peer-memory gradient exchange unavailable" in out, out[-3000:]
