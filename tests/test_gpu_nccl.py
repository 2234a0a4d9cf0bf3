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


def test_from_array_devices_matches_single_gpu():
    """The drop-in surface over two GPUs in ONE process (from_array(..., devices=[0, 1])): every sharded output equals the
    single-GPU object's, bit for bit, and the oracle's."""
    import numpy as np

    import oracle
    import pyflwdir_b200 as pfb
    from pyflwdir_b200 import _lib

    if _lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    z = oracle.synth_elevation(700, 520, seed=23)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.04)))
    one = pfb.from_array(d8, ftype="d8")
    two = pfb.from_array(d8, ftype="d8", devices=[0, 1])
    assert np.array_equal(two.idxs_ds, one.idxs_ds) and np.array_equal(two.idxs_pit, one.idxs_pit)
    assert np.array_equal(two.rank, one.rank)
    assert np.array_equal(two.upstream_area(), one.upstream_area()) and np.array_equal(two.basins(), one.basins())
    assert np.array_equal(two.stream_order(), one.stream_order())
    area = (np.abs(z) + np.float32(0.5)).astype(np.float32)
    assert np.array_equal(two.accuflux(area), one.accuflux(area))
    assert np.array_equal(two.upstream_area("km2"), one.upstream_area("km2"))
    drain = one.upstream_area() > 80
    assert np.array_equal(two.hand(drain, z), one.hand(drain, z))
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    assert np.array_equal(two.stream_order().ravel(), oracle.streams.strahler_order(ids, seq))
    assert np.array_equal(two.hand(drain, z).ravel(), oracle.dem.height_above_nearest_drain(ids, seq, drain.ravel(), z.ravel()))
    # unsharded entry points still work (devices[0])
    assert np.array_equal(two.idxs_seq, seq)


def test_fill_and_hand_on_the_second_device():
    """kernels that opt into large dynamic shared memory (the warp replay of dem.fill_depressions, phase A of the path-sum HAND) on
    device 1 after device 0 in the same process: the opt-in is per device"""
    import numpy as np

    import oracle
    import pyflwdir_b200 as pfb
    from pyflwdir_b200 import _lib, dem

    if _lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    flat = np.full((130, 140), 5.0, dtype=np.float32)  # one tie component of > 2048 cells: warp replay
    flat[0, :] = flat[-1, :] = flat[:, 0] = flat[:, -1] = 9.0
    flat[0, 70] = 1.0
    z = oracle.synth_elevation(300, 260, seed=5)
    d8 = oracle.synth_d8(z)
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    want_fill = oracle.dem.fill_depressions(flat)
    for device in (0, 1, 0):
        got = dem.fill_depressions(flat, device=device)
        assert np.array_equal(got[0], want_fill[0]) and np.array_equal(got[1], want_fill[1])
        flw = pfb.from_array(d8, ftype="d8", device=device)
        drain = flw.upstream_area() > 60
        assert np.array_equal(flw.hand(drain, z).ravel(), oracle.dem.height_above_nearest_drain(ids, seq, drain.ravel(), z.ravel()))
        assert flw._dev.info("hand_engine") == 1
