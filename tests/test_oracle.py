"""CPU tests: the plain-C oracle (oracle/sph_oracle.c) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.sh), and against the compiled reference itself when oracle/_ref is present."""
import numpy as np
import pytest

from conftest import golden, have_ref, run_ref
from compare import assert_close, rel_err
from opensph_b200 import abi
from oracle_port import OraclePort, build_lut

TOL = 1e-10      # north_star tolerance on derivatives
FLOOR = 1e-4     # per-quantity scale floor: lattice sums cancel, the reference's own solvers differ by more below it

DERIVS = ("acc", "du", "drho", "dS", "ddamage", "divv", "gradv", "corr", "vel")
EXACT = ("p", "cs", "reduce", "S", "pos")


def check_integrate(i, o, orc):
    assert np.array_equal(orc.a["ncnt"], o["ncnt"])
    for k in EXACT:
        if k in o and k in orc.a:
            assert_close(k, orc.a[k], o[k], 1e-14, FLOOR)
    for k in DERIVS:
        if k in o and k in orc.a:
            assert_close(k, orc.a[k], o[k], TOL, FLOOR)


def test_lut_matches_reference(lut):
    # LutKernel<3>(CubicSpline<3>) table, Kernel.h:85-101
    grad, val = build_lut(40000, 2.0)
    assert np.abs(grad - lut["lut_grad"]).max() <= 4e-16
    assert np.abs(val - lut["lut_val"]).max() <= 4e-16


@pytest.mark.parametrize("name", ["hello", "collision", "preset", "fluid", "gas"])
def test_integrate_matches_golden(name, lut):
    i, o = golden(f"{name}_in.snap"), golden(f"{name}_out.snap")
    orc = OraclePort(i, abi.setup_from_snapshot(i, lut))
    orc.integrate()
    check_integrate(i, o, orc)


@pytest.mark.parametrize("name", ["hello", "collision", "preset", "fluid", "gas"])
def test_neighbour_sets_bit_exact(name, lut):
    i, o = golden(f"{name}_in.snap"), golden(f"{name}_out.snap")
    orc = OraclePort(i, abi.setup_from_snapshot(i, lut))
    orc.integrate()  # clamps h like the reference before the finder is built
    off, idx = orc.neighbours()
    assert np.array_equal(off, o["nbr_offsets"])
    assert np.array_equal(idx, o["nbr_idx"])


def test_symmetric_solver_agrees(lut):
    # the reference's own cross-check: SymmetricSolver == AsymmetricSolver (solvers/test/Solvers.cpp:178-216)
    i, o = golden("hello_in.snap"), golden("hello_sym_out.snap")
    orc = OraclePort(i, abi.setup_from_snapshot(i, lut))
    orc.integrate()
    for k in ("acc", "du", "drho", "dS", "divv"):
        assert_close(k, orc.a[k], o[k], TOL, FLOOR)


@pytest.mark.parametrize("name,integrator", [("collision_pc3", "pc"), ("hello_pc3", "pc"), ("fluid_euler3", "euler"), ("gas_pc3", "pc")])
def test_time_steps_match_golden(name, integrator, lut):
    base = name.split("_")[0]
    i, o = golden(f"{base}_in.snap"), golden(f"{name}.snap")
    orc = OraclePort(i, abi.setup_from_snapshot(i, lut))
    consts = abi.run_constants(i)
    dts = o["dt_history"]
    orc.last_dt.value = consts["initial_dt"]
    assert dts[0] == consts["initial_dt"]
    for s in range(len(dts) - 1):
        step = orc.step_pc if integrator == "pc" else orc.step_euler
        dt, _ = step(float(dts[s]), consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1], (s, dt, dts[s + 1])
    for k in ("pos", "vel", "rho", "u", "S", "damage", "acc", "du", "drho", "dS"):
        if k in o and k in orc.a:
            assert_close(k, orc.a[k], o[k], 1e-9, FLOOR)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("args", [
    ["--config", "hello", "--n", 4000, "--solver", "asym", "--jitter", 11],
    ["--config", "collision_preset", "--n", 4000, "--jitter", 12],
    ["--config", "preset", "--n", 4000],
    ["--config", "preset_const_h", "--n", 3000, "--jitter", 2],
    ["--config", "fluid", "--n", 4000, "--jitter", 13],
    ["--config", "collision_preset", "--n", 3000, "--jitter", 14, "--enforcing"],
    ["--config", "collision_preset", "--n", 3000, "--jitter", 15, "--continuity-undamaged"],
    ["--config", "collision_preset", "--n", 3000, "--jitter", 16, "--sum-all", "--corrected", 0],
    ["--config", "hello", "--n", 3000, "--solver", "asym", "--jitter", 17, "--const-h", "--corrected", 1],
    ["--config", "collision_preset", "--n", 3000, "--jitter", 18, "--finder", "grid"],  # UniformGridFinder instead of KdTree
    ["--config", "gas", "--n", 3000, "--jitter", 19],                                     # IdealGasEos
])
def test_port_against_live_reference(args, tmp_path):
    i, o = run_ref(str(tmp_path), args)
    orc = OraclePort(i)
    orc.integrate()
    check_integrate(i, o, orc)
