"""The delta-SPH terms (core/sph/equations/DeltaSph.h, added by getStandardEquations with SPH_USE_DELTASPH,
StandardSets.cpp:64-67): the renormalised density gradient G_i = sum_j V_j (rho_j - rho_i) C_i grad W_ij, stored by every
evaluation; the density diffusion delta hbar cbar psi_ij . grad W_ij with psi built from the G of the PREVIOUS evaluation; the
velocity diffusion alpha hbar cbar pi_ij grad W_ij. Golden vectors come from the reference run with SPH_USE_DELTASPH
(tests/golden/make_golden.sh; delta = 0.1, alpha = 0.05 for the solid so that the terms are far above rounding, the library
defaults 0.01 for the fluid); CPU tests pin the oracle and the product's arithmetic, the -m gpu tests the device path through
the C ABI (all pair-kernel variants, one evaluation, PredictorCorrector steps where the stored gradient feeds the next
evaluation, solid and fluid)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from compare import assert_close
from opensph_b200 import abi
from oracle_port import OraclePort

FLOOR = 1e-4
STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag",
            "drho_grad")
OUT = ("acc", "du", "drho", "dS", "divv", "gradv", "corr", "drho_grad")


def golden(name):
    """Golden snapshot with the H component of the stored gradient cleared: without the correction tensor the reference
    leaves (h_i - h_j) * (dW/dq)/q sums there (the H lane of its SIMD kernel gradient), which nothing reads (dot() and the
    tensor product use x, y, z only); the drop-in keeps 0."""
    from conftest import golden as load
    snap = dict(load(name))
    if "drho_grad" in snap:
        snap["drho_grad"] = np.array(snap["drho_grad"], copy=True)
        snap["drho_grad"][:, 3] = 0.
    return snap


def test_oracle_deltasph_matches_golden(lut):
    i, o = golden("deltasph_in.snap"), golden("deltasph_out.snap")
    setup = abi.setup_from_snapshot(i, lut)
    assert setup.cfg.flags & abi.FLAG_DELTASPH and setup.deltasph_delta == 0.1 and setup.deltasph_alpha == 0.05
    orc = OraclePort(i, setup)
    orc.integrate()
    assert np.array_equal(orc.a["ncnt"], o["ncnt"])
    # the terms are not rounding-level quantities in this input: the run without them differs visibly
    plain = golden("collision_out.snap")
    assert np.abs(o["drho"] - plain["drho"]).max() > 1e-3 * np.abs(plain["drho"]).max()
    assert np.abs(o["acc"] - plain["acc"]).max() > 1e-4 * np.abs(plain["acc"]).max()
    assert np.abs(o["drho_grad"]).max() > 0.
    for k in OUT:
        assert_close(k, orc.a[k], o[k], 1e-10, FLOOR)


@pytest.mark.parametrize("case", ["deltasph", "deltasph_fluid"])
def test_oracle_deltasph_steps_match_golden(case, lut):
    """Three PredictorCorrector steps: the gradient stored by one evaluation enters the density diffusion of the next."""
    i = golden("deltasph_in.snap" if case == "deltasph" else "fluid_in.snap")
    o = golden(case + "_pc3.snap")
    setup = abi.setup_from_snapshot(o, lut)
    assert setup.cfg.flags & abi.FLAG_DELTASPH
    consts = abi.run_constants(o)
    orc = OraclePort(i, setup)
    orc.last_dt = C.c_double(consts["initial_dt"])
    dts = o["dt_history"]
    for s in range(len(dts) - 1):
        dt, _ = orc.step_pc(float(dts[s]), consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1]
    for k in ("pos", "vel", "rho", "u", "S", "drho_grad"):
        if k in o and k in orc.a:
            assert_close(k, orc.a[k], o[k], 1e-9, FLOOR)


@pytest.mark.parametrize("masked", [0, 1])
def test_product_deltasph_math_matches_golden(masked, lut):
    src = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
    lib = os.path.join(ROOT, "tests", "csrc", "libhostcheck.so")
    deps = [src, os.path.join(ROOT, "opensph_b200", "csrc", "sph_math.cuh"), os.path.join(ROOT, "opensph_b200", "csrc", "grav_math.cuh"),
            os.path.join(ROOT, "oracle", "sph_oracle.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", lib])
    hostcheck = C.CDLL(lib)
    i, o = golden("deltasph_in.snap"), golden("deltasph_out.snap")
    st = OraclePort(i, abi.setup_from_snapshot(i, lut))
    # a non-zero gradient of a "previous evaluation", so that the psi term is exercised by a single evaluation as well;
    # the oracle (pinned above and on the three-step golden) is the judge of that case
    rng = np.random.default_rng(5)
    g0 = np.zeros((st.n, 4))
    g0[:, :3] = rng.normal(size=(st.n, 3)) * np.abs(o["drho_grad"][:, :3]).max()
    ref = OraclePort(i, abi.setup_from_snapshot(i, lut))
    for port in (st, ref):
        port.a["drho_grad"][:] = g0
    ref.integrate()
    off = o["nbr_offsets"].astype(np.uint64)
    idx = o["nbr_idx"].astype(np.uint32)
    hostcheck.hostcheck_integrate(C.byref(st.state), C.byref(st.setup.cfg), st.setup.materials, C.c_uint32(st.setup.n_materials),
                                  off.ctypes.data_as(C.POINTER(C.c_uint64)), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int(masked))
    assert np.abs(ref.a["drho"] - o["drho"]).max() > 1e-3 * np.abs(o["drho"]).max()  # (the injected gradient matters)
    for k in ("acc", "du", "drho", "dS", "divv", "drho_grad"):
        assert_close(k, st.a[k], ref.a[k], 1e-10, FLOOR)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_gpu_deltasph_matches_golden(variant, lut):
    from opensph_b200.engine import Engine
    i, o = golden("deltasph_in.snap"), golden("deltasph_out.snap")
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
@pytest.mark.parametrize("variant", [0, 2])
def test_gpu_deltasph_with_a_previous_gradient_matches_oracle(variant, lut):
    """One evaluation on a state whose stored gradient is not zero (what every evaluation after the first one sees)."""
    from opensph_b200.engine import Engine
    i, o = golden("deltasph_in.snap"), golden("deltasph_out.snap")
    setup = abi.setup_from_snapshot(i, lut)
    rng = np.random.default_rng(5)
    state = {k: np.array(v, copy=True) for k, v in i.items()}
    state["drho_grad"] = np.zeros((len(i["mass"]), 4))
    state["drho_grad"][:, :3] = rng.normal(size=(len(i["mass"]), 3)) * np.abs(o["drho_grad"][:, :3]).max()
    ref = OraclePort(state, abi.setup_from_snapshot(i, lut))
    ref.integrate()
    with Engine(setup, len(i["mass"])) as eng:
        eng.set_variant(variant)
        eng.upload_state(state, STATE_IN)
        eng.integrate()
        got = eng.download_state(list(OUT) + ["ncnt"])
    assert np.array_equal(got["ncnt"], ref.a["ncnt"])
    for k in OUT:
        assert_close(k, got[k], ref.a[k], 1e-10, FLOOR)


@pytest.mark.gpu
@pytest.mark.parametrize("case,batched", [("deltasph", False), ("deltasph", True), ("deltasph_fluid", False)])
def test_gpu_deltasph_steps_match_golden(case, batched, lut):
    from opensph_b200.engine import Engine
    i = golden("deltasph_in.snap" if case == "deltasph" else "fluid_in.snap")
    o = golden(case + "_pc3.snap")
    setup = abi.setup_from_snapshot(o, lut)
    consts = abi.run_constants(o)
    dts = o["dt_history"]
    names = [k for k in STATE_IN + ("acc", "drho", "du", "dS", "ddamage") if k in i]
    with Engine(setup, len(i["mass"])) as eng:
        eng.upload_state(i, names)
        eng.set_last_timestep(consts["initial_dt"])
        if batched:
            hist, _, _ = eng.run_pc(len(dts) - 1, float(dts[0]), consts["max_dt"])
            assert np.allclose(hist, dts[1:], rtol=1e-9, atol=0)
        else:
            for s in range(len(dts) - 1):
                dt, _, _ = eng.step_pc(float(dts[s]), consts["max_dt"])
                assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1], (s, dt, dts[s + 1])
        got = eng.download_state([k for k in ("pos", "vel", "rho", "u", "S", "damage", "drho_grad") if k in o])
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)


@pytest.mark.gpu
def test_gpu_deltasph_is_rejected_where_it_is_not_implemented(lut):
    from opensph_b200.engine import Engine, SphGpuError
    i = golden("deltasph_in.snap")
    for extra in (abi.FLAG_BALSARA, abi.FLAG_XSPH):
        setup = abi.setup_from_snapshot(i, lut)
        setup.cfg.flags |= extra
        with pytest.raises(SphGpuError) as e:
            Engine(setup, len(i["mass"]))
        assert e.value.code == abi.E_INVALID
    setup = abi.setup_from_snapshot(i, lut)
    with Engine(setup, len(i["mass"])) as eng:
        with pytest.raises(SphGpuError):
            eng.set_variant(4)  # the symmetric formulation has no delta-SPH terms
