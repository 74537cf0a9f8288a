"""Time-frame sharding on real GPUs (needs >= 2): torchrun of tests/run_sharded_gpu.py; skipped on one GPU."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_time_sharded_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (covered on CPU by tests/test_sharding_gloo.py)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "run_sharded_gpu.py"), "30", "32"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
