"""Multi-GPU parity: rows sharded over 2 GPUs, NCCL allreduce inside the engine.
Skipped on single-GPU boxes; the host-side logic is covered on CPU by tests/test_sharding_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_solves_match_oracle():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "run_sharded_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(res.stdout[-3000:])
    print(res.stderr[-3000:])
    assert res.returncode == 0
    assert "MISMATCH" not in res.stdout
