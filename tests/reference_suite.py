"""Runs the REFERENCE'S OWN hot-path test files against pyflwdir_b200.

The files are not part of this repository: oracle/make_ref.sh places them (with tests/data/*.asc) in the git-ignored
oracle/_ref/tests/, which travels to the GPU box with the gpurun snapshot. Here `pyflwdir` is aliased to `pyflwdir_b200`
(package and submodules), then pytest runs the unmodified files. Needs a CUDA device (the package has no CPU fallback).

    python tests/reference_suite.py [--report profiles/reference_suite_rNN.md] [pytest args...]
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
# the reference test files that exercise the hot path (last row) + the codecs either side of it
FILES = ["test_core.py", "test_core_xx.py", "test_streams_basins.py", "test_pyflwdir.py", "test_flwdir.py", "test_dem.py",
         "test_basins.py", "test_arithmetics.py", "test_regions.py", "test_gis_utils.py"]
SUBMODULES = ["core", "core_d8", "core_ldd", "core_nextxy", "core_conversion", "streams", "basins", "dem", "arithmetics",
              "regions", "rivers", "gis_utils", "pyflwdir", "flwdir", "subgrid", "upscale"]


def alias():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    try:
        importlib.import_module("affine")
    except ImportError:  # the tests import affine.Affine: same stand-in the oracle uses
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
    import pyflwdir_b200

    sys.modules["pyflwdir"] = pyflwdir_b200
    missing = []
    for sub in SUBMODULES:
        try:
            sys.modules["pyflwdir." + sub] = importlib.import_module("pyflwdir_b200." + sub)
        except ImportError:
            missing.append(sub)
    return missing


class Collect:
    def __init__(self):
        self.rows = []

    def pytest_runtest_logreport(self, report):
        if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
            why = ""
            if report.outcome != "passed":
                why = str(report.longrepr).strip().splitlines()[-1][:160] if report.longrepr else ""
            outcome = report.outcome if report.when == "call" else ("error" if report.outcome == "failed" else report.outcome)
            self.rows.append((report.nodeid, outcome, why))

    def pytest_collectreport(self, report):
        if report.failed:
            why = str(report.longrepr).strip().splitlines()[-1][:160]
            self.rows.append((report.nodeid, "collection error", why))


def main(argv):
    import pytest

    report_path = None
    if "--report" in argv:
        i = argv.index("--report")
        report_path = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    if not os.path.isdir(REF_TESTS):
        print(f"reference tests not found under {REF_TESTS}: run oracle/make_ref.sh where /root/reference is mounted")
        return 2
    missing = alias()
    files = [os.path.join(REF_TESTS, f) for f in FILES if os.path.exists(os.path.join(REF_TESTS, f))]
    col = Collect()
    rc = pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", REF_TESTS, "-c", os.devnull, "--continue-on-collection-errors"]
                     + files + argv, plugins=[col])
    counts = {}
    for _, o, _ in col.rows:
        counts[o] = counts.get(o, 0) + 1
    lines = ["# The reference's own tests against pyflwdir_b200", "",
             "`python tests/reference_suite.py` on the GPU box: `pyflwdir` aliased to `pyflwdir_b200`, the UNMODIFIED files of "
             "`/root/reference/tests` (copied by `oracle/make_ref.sh` into the git-ignored `oracle/_ref/tests`).", "",
             "Totals: " + ", ".join(f"{k}: {v}" for k, v in sorted(counts.items())),
             "Submodules without a mirror in pyflwdir_b200 (out of scope, SURVEY.md section 2): " + (", ".join(missing) or "none"), "",
             "| test | outcome | reason (last line) |", "|---|---|---|"]
    for nodeid, o, why in col.rows:
        lines.append(f"| `{nodeid.split('tests/')[-1]}` | {o} | {why.replace('|', '/')} |")
    text = "\n".join(lines) + "\n"
    if report_path:
        os.makedirs(os.path.dirname(os.path.abspath(report_path)), exist_ok=True)
        with open(report_path, "w") as f:
            f.write(text)
    print(text[:3000])
    return int(rc)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
