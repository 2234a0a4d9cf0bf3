"""Row-tiled solve over real NCCL (one process per GPU). Needs >= 2 GPUs: skipped on single-GPU boxes, where the
same step functions are covered with emulated exchanges by tests/test_gpu_tiled.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_tiled_over_nccl(world, tmp_path):
    from pyflwdir_b200 import _lib

    if _lib.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_nccl_worker.py"), str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert f"NCCL_TILED_OK {world}" in out.stdout
    assert f"NCCL_SWEEPS_OK {world}" in out.stdout
    assert f"NCCL_ERROR_AGREEMENT_OK {world}" in out.stdout
