"""GPU (>= 2 devices on the box, skipped otherwise): the data-parallel sweeps across processes, one per GPU — the fused
GEMM -> reduce-scatter -> all-gather exchange over NVLink peer memory against NCCL and against the single-GPU full batch,
at a small shape and at the bench's own shard shape (rows = 8192 / N, D = 4096); aliasing cases of the exchange; C5.
tests/dp_check_multigpu.py is the per-rank program; this test launches it with torch.distributed.run."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    import torch
    return torch.cuda.device_count()


def _launch(n, *args):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_check_multigpu.py"), *[str(a) for a in args]]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.parametrize("B,D", [(2048, 1024), (8192, 4096)])
def test_data_parallel_sweeps_match_nccl_and_the_full_batch(B, D):
    n = _gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    res = _launch(n, B, D)
    assert res.returncode == 0, (res.stdout[-3000:], res.stderr[-3000:])
    assert res.stdout.count("C4 B=") == n, res.stdout[-2000:]
