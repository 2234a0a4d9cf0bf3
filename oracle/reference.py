"""Import the REAL reference (Deltares/pyflwdir, numba): from /root/reference when it is mounted (this container),
else from the git-ignored install under oracle/_ref/ (oracle/make_ref.sh), which travels to the GPU box with the
gpurun snapshot.

Used by tests/golden/make_golden.py (golden-vector generation), by CPU tests that pin the C oracle against the live
reference, and by bench.py's CPU legs (`--impl reference`, `cpu_baseline`). TEST / MEASUREMENT INFRASTRUCTURE ONLY:
nothing under pyflwdir_b200/ imports it."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PFD_REFERENCE_ROOT", "/root/reference")  # data files (tests/data, examples) live only here
_CANDIDATES = [os.environ.get("PFD_REFERENCE_ROOT", "/root/reference"), os.path.join(HERE, "_ref")]


def root():
    for r in _CANDIDATES:
        if r and os.path.isfile(os.path.join(r, "pyflwdir", "__init__.py")):
            return r
    return None


def available():
    return root() is not None


def load():
    """Returns the reference `pyflwdir` module (numba JIT cache redirected to a writable dir)."""
    r = root()
    if r is None:
        raise ImportError("reference not found (neither /root/reference nor oracle/_ref; run oracle/make_ref.sh)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/pfd_numba_cache")
    try:
        importlib.import_module("affine")
    except ImportError:
        sys.path.insert(0, os.path.join(HERE, "_stubs"))
    if r not in sys.path:
        sys.path.insert(0, r)
    return importlib.import_module("pyflwdir")
