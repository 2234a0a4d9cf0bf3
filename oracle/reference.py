"""Import the REAL reference (Deltares/pyflwdir, numba) when it is mounted at /root/reference.

Used only by tests/golden/make_golden.py (golden-vector generation) and by CPU tests that pin the C oracle
against the live reference. /root/reference does not exist on the GPU box: everything that runs there uses the
committed golden vectors instead. TEST INFRASTRUCTURE ONLY."""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("PFD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyflwdir"))


def load():
    """Returns the reference `pyflwdir` module (numba JIT cache redirected to a writable dir)."""
    if not available():
        raise ImportError("reference not mounted")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/pfd_numba_cache")
    stubs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")
    try:
        importlib.import_module("affine")
    except ImportError:
        sys.path.insert(0, stubs)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("pyflwdir")
