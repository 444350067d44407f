"""Partitioned run vs single-GPU run on real GPUs (needs >= 2 devices; tests/dist_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,no_p2p", [(8, "0"), (8, "1"), (20, "0")])
def test_partitioned_step_matches_single_gpu(gpu_ctx, n, no_p2p):
    """n = 8: Jacobi-CG pressure solve; n = 20: 9261 pressure unknowns -> per-rank AMG (additive Schwarz).
    no_p2p = 1: NCCL transport instead of the peer-memory kernels."""
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs on the box (run under `gpurun --gpus 2`); the host-side logic is covered by tests/test_parallel.py")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join(ROOT, "tests", "dist_check.py"), str(n)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1", FB_NO_P2P=no_p2p))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" OK") == 2
    if no_p2p == "1":
        assert "transport: NCCL" in out.stdout, out.stdout[-2000:]
