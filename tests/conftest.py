import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_STRICT = os.path.join(ROOT, "oracle", "_ref", "sph_ref_strict")
REF_FAST = os.path.join(ROOT, "oracle", "_ref", "sph_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def golden(name):
    from opensph_b200.snapshot import read_snapshot
    return read_snapshot(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def lut():
    return golden("lut.snap")


def have_ref():
    return os.path.exists(REF_STRICT)


def run_ref(tmpdir, args, binary=None):
    """Runs the compiled reference (oracle/_ref) and returns (in_snapshot, out_snapshot)."""
    from opensph_b200.snapshot import read_snapshot
    binary = binary or REF_STRICT
    fin, fout = os.path.join(tmpdir, "in.snap"), os.path.join(tmpdir, "out.snap")
    subprocess.check_call([binary, "snapshot", "--in", fin, "--out", fout] + [str(a) for a in args],
                          stdout=subprocess.DEVNULL)
    return read_snapshot(fin), read_snapshot(fout)
