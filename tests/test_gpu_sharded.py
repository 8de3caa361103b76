"""Multi-GPU path on real devices: window sharding + NCCL gather must reproduce the single-GPU result exactly.
Needs >= 2 GPUs on the box (skipped otherwise); the plumbing itself is covered on CPU by tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_equals_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "check_sharded.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(p.stdout[-2000:])
    assert p.returncode == 0, p.stderr[-3000:]
    assert "SHARDED_CHECK OK" in p.stdout
