"""The XSph term (core/sph/equations/XSph.h): the velocity correction eps * sum_j m_j (v_j - v_i) W_ij / rhobar,
taken out of the velocities before the derivatives are evaluated and put back afterwards. Golden vectors come from the
reference run with SPH_USE_XSPH (tests/golden/make_golden.sh); CPU tests pin the oracle and the product's arithmetic, the
-m gpu tests the device path through the C ABI (all pair-kernel variants, single evaluation and PredictorCorrector steps)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from compare import assert_close
from opensph_b200 import abi
from oracle_port import OraclePort

FLOOR = 1e-4
STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag", "xsph")
OUT = ("acc", "du", "drho", "dS", "divv", "gradv", "corr", "vel", "xsph")


def test_oracle_xsph_matches_golden(lut):
    i, o = golden("xsph_in.snap"), golden("xsph_out.snap")
    setup = abi.setup_from_snapshot(i, lut)
    assert setup.cfg.flags & abi.FLAG_XSPH and setup.xsph_eps == 0.5
    orc = OraclePort(i, setup)
    orc.integrate()
    assert np.array_equal(orc.a["ncnt"], o["ncnt"])
    assert np.abs(o["xsph"]).max() > 1.0  # the correction is not a rounding-level quantity in this input
    for k in OUT:
        assert_close(k, orc.a[k], o[k], 1e-10, FLOOR)


def test_oracle_xsph_steps_match_golden(lut):
    i, o = golden("xsph_in.snap"), golden("xsph_pc3.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    orc = OraclePort(i, setup)
    orc.last_dt = C.c_double(consts["initial_dt"])
    dts = o["dt_history"]
    for s in range(len(dts) - 1):
        dt, _ = orc.step_pc(float(dts[s]), consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1]
    for k in ("pos", "vel", "rho", "u", "S", "xsph"):
        assert_close(k, orc.a[k], o[k], 1e-9, FLOOR)


@pytest.mark.parametrize("masked", [0, 1])
def test_product_xsph_math_matches_golden(masked, lut):
    src = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
    lib = os.path.join(ROOT, "tests", "csrc", "libhostcheck.so")
    deps = [src, os.path.join(ROOT, "opensph_b200", "csrc", "sph_math.cuh"), os.path.join(ROOT, "opensph_b200", "csrc", "grav_math.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", lib])
    hostcheck = C.CDLL(lib)
    i, o = golden("xsph_in.snap"), golden("xsph_out.snap")
    st = OraclePort(i, abi.setup_from_snapshot(i, lut))
    off = o["nbr_offsets"].astype(np.uint64)
    idx = o["nbr_idx"].astype(np.uint32)
    hostcheck.hostcheck_integrate(C.byref(st.state), C.byref(st.setup.cfg), st.setup.materials, C.c_uint32(st.setup.n_materials),
                                  off.ctypes.data_as(C.POINTER(C.c_uint64)), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int(masked))
    for k in ("acc", "du", "drho", "dS", "divv", "vel", "xsph"):
        assert_close(k, st.a[k], o[k], 1e-10, FLOOR)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_gpu_xsph_matches_golden(variant, lut):
    from opensph_b200.engine import Engine
    i, o = golden("xsph_in.snap"), golden("xsph_out.snap")
    setup = abi.setup_from_snapshot(i, lut)
    with Engine(setup, len(i["mass"])) as eng:
        eng.set_variant(variant)
        eng.upload_state(i, STATE_IN)
        eng.integrate()
        got = eng.download_state(list(OUT) + ["ncnt"])
    assert np.array_equal(got["ncnt"], o["ncnt"])
    for k in OUT:
        assert_close(k, got[k], o[k], 1e-10, FLOOR)


@pytest.mark.gpu
def test_gpu_xsph_steps_match_golden(lut):
    from opensph_b200.engine import Engine
    i, o = golden("xsph_in.snap"), golden("xsph_pc3.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    dts = o["dt_history"]
    with Engine(setup, len(i["mass"])) as eng:
        eng.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
        eng.set_last_timestep(consts["initial_dt"])
        for s in range(len(dts) - 1):
            dt, _, _ = eng.step_pc(float(dts[s]), consts["max_dt"])
            assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1], (s, dt, dts[s + 1])
        got = eng.download_state(["pos", "vel", "rho", "u", "S", "damage", "xsph"])
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)


@pytest.mark.gpu
def test_gpu_xsph_is_rejected_where_it_is_not_implemented(lut):
    from opensph_b200.engine import Engine, SphGpuError
    i = golden("xsph_in.snap")
    setup = abi.setup_from_snapshot(i, lut)
    setup.cfg.flags |= abi.FLAG_BALSARA
    with pytest.raises(SphGpuError) as e:
        Engine(setup, len(i["mass"]))
    assert e.value.code == abi.E_INVALID


@pytest.mark.gpu
def test_gpu_xsph_batched_steps_match_golden(lut):
    """The same three steps queued back to back (sphgpu_run_pc: fused corrector + predictor, time step fed back on the device)."""
    from opensph_b200.engine import Engine
    i, o = golden("xsph_in.snap"), golden("xsph_pc3.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    dts = o["dt_history"]
    with Engine(setup, len(i["mass"])) as eng:
        eng.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
        eng.set_last_timestep(consts["initial_dt"])
        hist, _, _ = eng.run_pc(len(dts) - 1, float(dts[0]), consts["max_dt"])
        got = eng.download_state(["pos", "vel", "rho", "u", "S", "damage", "xsph"])
    assert np.allclose(hist, dts[1:], rtol=1e-9, atol=0)
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)
