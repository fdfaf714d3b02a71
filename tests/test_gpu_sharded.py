"""Multi-GPU parity: needs >= 2 GPUs on the box, skipped otherwise.  The host logic
of the sharding is covered on CPU by tests/test_sharded_gloo.py."""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_forms_equal_single_gpu():
    n = min(torch.cuda.device_count(), 4)
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sharded_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", worker]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    print(res.stdout.strip().splitlines()[-1])
