"""run_mcmc(cluster="torch") over NCCL on two real GPUs (skipped on a one-GPU box; the sharding / gather
logic itself is covered on CPU by tests/test_multi_rank_gloo.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_run_equals_one_gpu_run():
    import drjacoby_b200 as drj
    if drj.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK full" in out.stdout and "OK summary" in out.stdout and "OK few chains" in out.stdout
