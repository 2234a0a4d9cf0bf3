"""CPU test of the N > 1 path of bench.py: two processes, gloo backend, rendezvous on 127.0.0.1."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_plumbing():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DIST_OK" in out.stdout


def test_reference_arm_nonzero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_json_contract():
    """bench.py --impl reference on a tiny raster (CPU only here): one JSON line with the contract's keys."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "256",
                          "--cpu-size", "256", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Mcells/s" and line["value"] > 0
    # the real numba reference when it is installed (oracle/make_ref.sh or /root/reference), else the C port
    from oracle import reference

    assert line["cpu_baseline"]["kind"] == ("reference" if reference.available() else "port")
    assert line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"] and "cpu_sample" in line["config"]


def test_reference_arm_port_fallback():
    """Without the reference install the CPU arm times the oracle port and says so."""
    import json

    env = dict(os.environ, PFD_BENCH_CPU="port")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "256",
                          "--cpu-size", "256", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                         cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["kind"] == "port"
