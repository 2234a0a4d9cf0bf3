#!/usr/bin/env bash
# Installs the UNMODIFIED reference (Deltares/pyflwdir, pure Python + numba) into the git-ignored oracle/_ref/ so
# that it travels to the GPU box with the gpurun snapshot (like the built .so files) and bench.py can time the
# reference's own numba CPU path there (cpu_baseline.kind = "reference").
#
# The reference is a flit pure-python package. `pip install --no-index --target oracle/_ref /root/reference` fails
# in this image because the build backend (flit_core) is not in the offline wheelhouse; what that install would do
# for a pure-python wheel is exactly this: place the package directory on the target path. Nothing is copied into
# the git history (oracle/_ref/ is in .gitignore), nothing of it is imported by the product (tests/test_abi.py).
# The missing `affine` dependency is covered by oracle/_stubs/affine.py (see oracle/reference.py).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${PFD_REFERENCE_ROOT:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/pyflwdir" ]; then
    echo "make_ref.sh: $SRC/pyflwdir not found (GPU box: uses the prebuilt oracle/_ref)" >&2
    exit 0
fi
if python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target "$DST" "$SRC" 2>/dev/null; then
    echo "make_ref.sh: pip-installed the reference into $DST"
else
    rm -rf "$DST/pyflwdir"
    mkdir -p "$DST/pyflwdir"
    cp "$SRC"/pyflwdir/*.py "$DST/pyflwdir/"
    echo "make_ref.sh: pip install unavailable (no flit_core offline); placed the package directory in $DST"
fi
# the reference's own test files + fixtures (run against pyflwdir_b200 by tests/reference_suite.py on the GPU box)
rm -rf "$DST/tests"
mkdir -p "$DST/tests/data"
cp "$SRC"/tests/*.py "$DST/tests/"
cp "$SRC"/tests/data/* "$DST/tests/data/"
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/REFERENCE_COMMIT" || true
