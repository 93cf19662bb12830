"""CPU check of the PRODUCT's per-pair / per-particle arithmetic (opensph_b200/csrc/sph_math.cuh, the functions the
CUDA kernels call) through a host harness, against golden vectors of the reference. The kernels' indexing is
covered by the -m gpu tests; this pins the formulas (incl. the single-pass correction-tensor factorisation)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from compare import assert_close
from opensph_b200 import abi
from oracle_port import OraclePort

SRC = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
LIB = os.path.join(ROOT, "tests", "csrc", "libhostcheck.so")


@pytest.fixture(scope="module")
def hostcheck():
    deps = [SRC, os.path.join(ROOT, "opensph_b200", "csrc", "sph_math.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", SRC, "-o", LIB])
    return C.CDLL(LIB)


@pytest.mark.parametrize("masked", [0, 1])
@pytest.mark.parametrize("name", ["hello", "collision", "preset", "fluid", "gas"])
def test_product_math_matches_golden(name, masked, hostcheck, lut):
    i, o = golden(f"{name}_in.snap"), golden(f"{name}_out.snap")
    st = OraclePort(i, abi.setup_from_snapshot(i, lut))
    off = o["nbr_offsets"].astype(np.uint64)
    idx = o["nbr_idx"].astype(np.uint32)
    hostcheck.hostcheck_integrate(C.byref(st.state), C.byref(st.setup.cfg), st.setup.materials,
                                  C.c_uint32(st.setup.n_materials), off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                  idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int(masked))
    assert np.array_equal(st.a["ncnt"], o["ncnt"])
    for k in ("p", "cs", "reduce", "S"):
        if k in o and k in st.a:
            assert_close(k, st.a[k], o[k], 1e-14, 1e-4)
    for k in ("acc", "du", "drho", "dS", "ddamage", "divv", "gradv", "corr", "vel"):
        if k in o and k in st.a:
            assert_close(k, st.a[k], o[k], 1e-10, 1e-4)
