"""The artificial stress (core/sph/equations/av/Stress.cpp, added by getStandardEquations with SPH_AV_USE_STRESS,
StandardSets.cpp:72-74): every evaluation turns the total stress S - p I of a particle into as = -(S - p I)+ (the tensile principal
stresses, negated; StressAV::initialize) and the pairs add Pi_ij = xi (W_ij / W(h_i,h_i))^n (as_i / rho_i^2 + as_j / rho_j^2) to the
acceleration and the heating. Golden vectors come from the reference run with SPH_AV_USE_STRESS (tests/golden/make_golden.sh;
factor 0.2, 150 of the 537 particles in tension); CPU tests pin the oracle and the product's arithmetic, the -m gpu tests the
device path through the C ABI (all pair-kernel variants, one evaluation, PredictorCorrector steps)."""
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
STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag", "wp")
OUT = ("acc", "du", "drho", "dS", "divv", "gradv", "corr", "av_stress")


def test_oracle_stress_av_matches_golden(lut):
    i, o = golden("stressav_in.snap"), golden("stressav_out.snap")
    setup = abi.setup_from_snapshot(i, lut)
    assert setup.cfg.flags & abi.FLAG_STRESS_AV and setup.stress_av_exponent == 4.0 and setup.stress_av_factor == 0.2
    orc = OraclePort(i, setup)
    orc.integrate()
    assert np.array_equal(orc.a["ncnt"], o["ncnt"])
    # the term is not a rounding-level quantity in this input: the run without it differs visibly
    plain = golden("collision_out.snap")
    assert np.abs(o["acc"] - plain["acc"]).max() > 1e-3 * np.abs(plain["acc"]).max()
    assert np.abs(o["du"] - plain["du"]).max() > 1e-4 * np.abs(plain["du"]).max()
    assert (np.abs(o["av_stress"]).max(axis=1) > 0).sum() > 100
    for k in OUT:
        assert_close(k, orc.a[k], o[k], 1e-10, FLOOR)


def test_oracle_stress_av_steps_match_golden(lut):
    i, o = golden("stressav_in.snap"), golden("stressav_pc3.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    orc = OraclePort(i, setup)
    orc.last_dt = C.c_double(consts["initial_dt"])
    dts = o["dt_history"]
    for s in range(len(dts) - 1):
        dt, _ = orc.step_pc(float(dts[s]), consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1]
    for k in ("pos", "vel", "rho", "u", "S", "av_stress"):
        assert_close(k, orc.a[k], o[k], 1e-9, FLOOR)


@pytest.mark.parametrize("masked", [0, 1])
def test_product_stress_av_math_matches_golden(masked, lut):
    src = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
    lib = os.path.join(ROOT, "tests", "csrc", "libhostcheck.so")
    deps = [src, os.path.join(ROOT, "opensph_b200", "csrc", "sph_math.cuh"), os.path.join(ROOT, "opensph_b200", "csrc", "grav_math.cuh"),
            os.path.join(ROOT, "oracle", "sph_oracle.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", lib])
    hostcheck = C.CDLL(lib)
    i, o = golden("stressav_in.snap"), golden("stressav_out.snap")
    st = OraclePort(i, abi.setup_from_snapshot(i, lut))
    off = o["nbr_offsets"].astype(np.uint64)
    idx = o["nbr_idx"].astype(np.uint32)
    hostcheck.hostcheck_integrate(C.byref(st.state), C.byref(st.setup.cfg), st.setup.materials, C.c_uint32(st.setup.n_materials),
                                  off.ctypes.data_as(C.POINTER(C.c_uint64)), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int(masked))
    for k in ("acc", "du", "drho", "dS", "divv", "av_stress"):
        assert_close(k, st.a[k], o[k], 1e-10, FLOOR)


def test_product_av_stress_is_minus_the_positive_part_of_the_tensor():
    """avStressOf (Jacobi rotations) against numpy's eigh on random, degenerate, diagonal, rank-one, zero and badly scaled
    tensors: as = -V max(Lambda, 0) V^T is a function of the tensor, whatever the eigen-solver."""
    src = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
    lib = os.path.join(ROOT, "tests", "csrc", "libhostcheck.so")
    deps = [src, os.path.join(ROOT, "opensph_b200", "csrc", "sph_math.cuh"), os.path.join(ROOT, "opensph_b200", "csrc", "grav_math.cuh"),
            os.path.join(ROOT, "oracle", "sph_oracle.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", lib])
    hostcheck = C.CDLL(lib)
    rng = np.random.default_rng(21)
    mats = []
    for _ in range(2000):
        a = rng.normal(size=(3, 3)) * 10.0 ** rng.uniform(-3, 9)
        mats.append(0.5 * (a + a.T))
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    for lam in ([1., 1., 1.], [2., 2., -1.], [3., -1., -1.], [0., 0., 0.], [1e9, 1e-9, -1e9], [5., 0., 0.], [-1., -2., -3.]):
        mats.append(q @ np.diag(lam) @ q.T)          # degenerate / semi-definite / definite spectra in a rotated frame
        mats.append(np.diag(lam))                    # and already diagonal
    v = rng.normal(size=3)
    mats.append(np.outer(v, v))                      # rank one
    m = np.array(mats)
    sig = np.ascontiguousarray(np.stack([m[:, 0, 0], m[:, 1, 1], m[:, 2, 2], m[:, 0, 1], m[:, 0, 2], m[:, 1, 2]], axis=1))
    out = np.zeros_like(sig)
    hostcheck.hostcheck_av_stress(C.c_uint32(len(sig)), sig.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    w, vec = np.linalg.eigh(m)
    ref = -np.einsum("nik,nk,njk->nij", vec, np.maximum(w, 0.), vec)
    want = np.stack([ref[:, 0, 0], ref[:, 1, 1], ref[:, 2, 2], ref[:, 0, 1], ref[:, 0, 2], ref[:, 1, 2]], axis=1)
    scale = np.abs(sig).max(axis=1, keepdims=True) + 1e-300
    assert np.abs(out - want).max(initial=0.) <= 0 or (np.abs(out - want) / scale).max() <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_gpu_stress_av_matches_golden(variant, lut):
    from opensph_b200.engine import Engine
    i, o = golden("stressav_in.snap"), golden("stressav_out.snap")
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
@pytest.mark.parametrize("exponent", [4.0, 2.5])
def test_gpu_stress_av_exponents_match_oracle(exponent, lut):
    """The weighting function with the default integer exponent (repeated squaring on the device) and a fractional one (pow)."""
    from opensph_b200.engine import Engine
    i = golden("stressav_in.snap")
    setup = abi.setup_from_snapshot(i, lut)
    setup.stress_av_exponent, setup.stress_av_factor = exponent, 0.07
    ref = OraclePort(i, setup)
    ref.integrate()
    with Engine(setup, len(i["mass"])) as eng:
        eng.upload_state(i, STATE_IN)
        eng.integrate()
        got = eng.download_state(list(OUT) + ["ncnt"])
    assert np.array_equal(got["ncnt"], ref.a["ncnt"])
    for k in OUT:
        assert_close(k, got[k], ref.a[k], 1e-10, FLOOR)


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
def test_gpu_stress_av_steps_match_golden(batched, lut):
    from opensph_b200.engine import Engine
    i, o = golden("stressav_in.snap"), golden("stressav_pc3.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    dts = o["dt_history"]
    with Engine(setup, len(i["mass"])) as eng:
        eng.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
        eng.set_last_timestep(consts["initial_dt"])
        if batched:
            hist, _, _ = eng.run_pc(len(dts) - 1, float(dts[0]), consts["max_dt"])
            assert np.allclose(hist, dts[1:], rtol=1e-9, atol=0)
        else:
            for s in range(len(dts) - 1):
                dt, _, _ = eng.step_pc(float(dts[s]), consts["max_dt"])
                assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1], (s, dt, dts[s + 1])
        got = eng.download_state(["pos", "vel", "rho", "u", "S", "damage", "av_stress"])
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)


@pytest.mark.gpu
def test_gpu_stress_av_is_rejected_where_it_is_not_implemented(lut):
    from opensph_b200.engine import Engine, SphGpuError
    i = golden("stressav_in.snap")
    for extra in (abi.FLAG_BALSARA, abi.FLAG_XSPH, abi.FLAG_DELTASPH):
        setup = abi.setup_from_snapshot(i, lut)
        setup.cfg.flags |= extra
        with pytest.raises(SphGpuError) as e:
            Engine(setup, len(i["mass"]))
        assert e.value.code == abi.E_INVALID
    f = golden("fluid_in.snap")
    setup = abi.setup_from_snapshot(f, lut)
    setup.cfg.flags |= abi.FLAG_STRESS_AV  # no deviatoric stress to build the artificial stress from
    with pytest.raises(SphGpuError):
        Engine(setup, len(f["mass"]))
    setup = abi.setup_from_snapshot(i, lut)
    with Engine(setup, len(i["mass"])) as eng:
        with pytest.raises(SphGpuError):
            eng.set_variant(4)
