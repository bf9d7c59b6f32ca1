"""GPU, N >= 2: X-decomposed run over NCCL against the reference's single-rank golden trajectory
(the reference's own multi-rank criterion is nproc-independence, tests/integrated/test-delp2/runtest:13,
test-laplace/runtest:34).  Skipped on a one-GPU box; tests/test_mesh_gloo.py covers the host logic on CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 4, 8])
def test_hw3d_x_decomposed(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(29700 + n), os.path.join(ROOT, "tests", "multigpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
