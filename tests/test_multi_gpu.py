"""Multi-GPU correctness under pytest (needs >= 2 GPUs; skipped otherwise): one process per GPU through torchrun.
The worker (tools/test_multi_gpu.py) checks that the NVLink peer-mailbox all-reduce of the 7 floats (ehb_allreduce7, and
its fused form inside pose_backward / adam) equals NCCL's bit for bit across ranks, that a pose solve with the views
sharded over the ranks reproduces the single-rank solve, and that space-exploration scores sharded over the ranks and
all-gathered equal the single-rank scores exactly."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world", [2])
def test_two_ranks_allreduce_sharded_solve_and_exploration(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "test_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU checks ok" in r.stdout, r.stdout[-3000:]
