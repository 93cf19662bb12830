"""The symmetric formulation on the device (variant 4 of the pair stage): particles ranked by
smoothing length (makeRankH), every pair evaluated once by its particle of higher rank (findLowerRank) and added to both
(evalSymmetric, NeighborCountTerm) -- against the golden vector of the reference's own SymmetricSolver, against the
asymmetric golden vectors (the reference's cross-check, core/sph/solvers/test/Solvers.cpp:178-216), on live reference runs
with jittered smoothing lengths (so that the ranks matter), over PredictorCorrector steps, and with ghosts and giants."""
import numpy as np
import pytest

from conftest import golden, have_ref, run_ref
from compare import assert_close
from opensph_b200 import abi

pytestmark = pytest.mark.gpu
TOL, FLOOR = 1e-10, 1e-4
STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag")
DERIVS = ("acc", "du", "drho", "dS", "divv", "gradv")


def _integrate(i, setup, variant):
    from opensph_b200.engine import Engine
    eng = Engine(setup, len(i["mass"]))
    eng.set_variant(variant)
    eng.upload_state(i, STATE_IN)
    st = eng.integrate()
    return eng, st


def test_symmetric_variant_matches_symmetric_solver_golden(lut):
    i, o = golden("hello_in.snap"), golden("hello_sym_out.snap")  # the library-default SymmetricSolver of the reference
    eng, st = _integrate(i, abi.setup_from_snapshot(i, lut), 4)
    got = eng.download_state(list(DERIVS) + ["ncnt", "vel"])
    eng.close()
    assert np.array_equal(got["ncnt"], golden("hello_out.snap")["ncnt"])  # NeighborCountTerm: ++cnt_i, ++cnt_j
    assert st.pair_count == int(got["ncnt"].astype(np.int64).sum())
    for k in DERIVS + ("vel",):
        if k in o:
            assert_close(k, got[k], o[k], TOL, FLOOR)


@pytest.mark.parametrize("name", ["hello", "fluid", "gas"])
def test_symmetric_variant_matches_asymmetric_golden(name, lut):
    i, o = golden(f"{name}_in.snap"), golden(f"{name}_out.snap")
    eng, st = _integrate(i, abi.setup_from_snapshot(i, lut), 4)
    got = eng.download_state([k for k in DERIVS if k in o] + ["ncnt"])
    eng.close()
    assert np.array_equal(got["ncnt"], o["ncnt"])
    for k in got:
        if k != "ncnt":
            assert_close(k, got[k], o[k], TOL, FLOOR)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args", [["--config", "hello", "--n", 30000, "--solver", "sym", "--jitter", 5],
                                  ["--config", "collision_preset", "--n", 20000, "--solver", "sym", "--corrected", 0, "--jitter", 9]])
def test_symmetric_variant_against_live_symmetric_solver(args, tmp_path):
    i, o = run_ref(str(tmp_path), args)
    assert abi.run_constants(i)["solver"] == 0  # SolverEnum::SYMMETRIC_SOLVER
    eng, st = _integrate(i, abi.setup_from_snapshot(i), 4)
    got = eng.download_state([k for k in DERIVS if k in o] + ["ncnt"])
    # and three steps of the device integrator with the symmetric kernels keep working (lists are not used by them)
    eng.run_pc(3, 1e-4, 1e-3)
    eng.close()
    assert np.array_equal(got["ncnt"], o["ncnt"])
    for k in got:
        if k != "ncnt":
            assert_close(k, got[k], o[k], TOL, FLOOR)


def test_symmetric_variant_with_ghosts_and_giants(lut):
    """Ghost particles of higher rank own pairs with owned partners; a giant particle is paired by the two-level kernels."""
    from opensph_b200.engine import Engine
    i, o = golden("hello_in.snap"), golden("hello_out.snap")
    n = len(i["mass"])
    owned = np.where(i["pos"][:, 0] < 0)[0]
    ghost = np.where(i["pos"][:, 0] >= 0)[0]
    perm = np.concatenate([owned, ghost])
    part = {k: (v[perm] if (hasattr(v, "shape") and v.shape[:1] == (n,)) else v) for k, v in i.items()}
    setup = abi.setup_from_snapshot(i, lut)
    setup.materials[0].begin, setup.materials[0].end = 0, len(owned)
    with Engine(setup, len(owned), capacity=n) as eng:
        eng.set_variant(4)
        eng.upload_state({k: v[: len(owned)] for k, v in part.items() if k in STATE_IN}, STATE_IN)
        eng.upload_state({k: v[len(owned):] for k, v in part.items() if k in STATE_IN}, STATE_IN, first=len(owned))
        eng.upload("MATERIAL_ID", 0, np.zeros(len(ghost), np.uint32), first=len(owned))
        eng.set_active(n)
        eng.integrate()
        got = eng.download_state(["acc", "du", "drho", "divv", "ncnt"])
    assert np.array_equal(got["ncnt"], o["ncnt"][owned])
    for k in ("acc", "du", "drho", "divv"):
        assert_close(k, got[k], o[k][owned], TOL, FLOOR)
    # one giant smoothing length: symmetric (variant 4) against asymmetric (variant 0) on the same input
    big = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in i.items()}
    big["pos"][n // 3, 3] *= 6.0
    res = {}
    for variant in (0, 4):
        eng, st = _integrate(big, abi.setup_from_snapshot(i, lut), variant)
        res[variant] = eng.download_state(["acc", "du", "drho", "divv", "ncnt"])
        eng.close()
    assert np.array_equal(res[0]["ncnt"], res[4]["ncnt"]) and res[0]["ncnt"].max() > 2 * o["ncnt"].max()
    for k in ("acc", "du", "drho", "divv"):
        assert_close(k, res[4][k], res[0][k], TOL, FLOOR)


def test_symmetric_variant_rejects_what_the_symmetric_solver_rejects(lut):
    from opensph_b200.engine import Engine, SphGpuError
    i = golden("collision_in.snap")  # strain-rate correction tensor on: SymmetricSolver throws InvalidSetup (SymmetricSolver.cpp:41-44)
    with Engine(abi.setup_from_snapshot(i, lut), len(i["mass"])) as eng:
        with pytest.raises(SphGpuError) as e:
            eng.set_variant(4)
        assert e.value.code == abi.E_INVALID
