"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path through the C ABI against (a) the committed
golden vectors of the unmodified reference, (b) the plain-C oracle on seeded inputs, (c) the compiled reference itself
(oracle/_ref, travels with the snapshot) at larger sizes.

Bar (BASELINE.json north_star): neighbour lists bit-exact as sets; dv, drho, du, dS (and divv, gradv, C, dh/dt) within
1e-10 relative, normalised per quantity with a floor of 1e-4 * max|q| (lattice sums cancel; see tests/test_oracle.py)."""
import numpy as np
import pytest

from conftest import golden, have_ref, run_ref
from compare import assert_close
from opensph_b200 import abi
from opensph_b200.engine import Engine
from oracle_port import OraclePort

pytestmark = pytest.mark.gpu

TOL, FLOOR = 1e-10, 1e-4
DERIVS = ("acc", "du", "drho", "dS", "ddamage", "divv", "gradv", "corr", "vel")
EXACT = ("p", "cs", "reduce", "S", "pos")
STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag")


def gpu_integrate(snap, setup, variant=0):
    n = len(snap["mass"])
    eng = Engine(setup, n)
    eng.set_variant(variant)
    eng.upload_state(snap, STATE_IN)
    stats = eng.integrate()
    return eng, stats


def check_against(eng, stats, ref, names_exact=EXACT, names_derivs=DERIVS, tol=TOL, tol_exact=1e-12):
    got = eng.download_state([k for k in names_exact + names_derivs + ("ncnt",) if k in ref])
    assert np.array_equal(got["ncnt"], ref["ncnt"]), "NEIGHBOR_CNT differs"
    assert stats.neigh_min == ref["ncnt"].min() and stats.neigh_max == ref["ncnt"].max()
    assert stats.pair_count == int(ref["ncnt"].astype(np.int64).sum())
    for k in names_exact:
        if k in ref:
            assert_close(k, got[k], ref[k], tol_exact, FLOOR)  # (1 - D^3) cancels for D -> 1: a few hundred ulp
    for k in names_derivs:
        if k in ref:
            if k == "acc":
                assert np.all(got[k][:, 3] == 0.0)
            assert_close(k, got[k], ref[k], tol, FLOOR)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("name", ["hello", "collision", "preset", "fluid", "gas"])
def test_integrate_matches_golden(name, variant, lut):
    i, o = golden(f"{name}_in.snap"), golden(f"{name}_out.snap")
    eng, stats = gpu_integrate(i, abi.setup_from_snapshot(i, lut), variant)
    check_against(eng, stats, o)
    eng.close()


@pytest.mark.parametrize("name", ["hello", "collision", "preset", "fluid", "gas"])
def test_neighbour_lists_bit_exact(name, lut):
    i, o = golden(f"{name}_in.snap"), golden(f"{name}_out.snap")
    eng, _ = gpu_integrate(i, abi.setup_from_snapshot(i, lut))
    off, idx = eng.neighbours()
    assert np.array_equal(off, o["nbr_offsets"])
    assert np.array_equal(idx, o["nbr_idx"])
    eng.close()


@pytest.mark.parametrize("name,integrator", [("collision_pc3", "pc"), ("hello_pc3", "pc"), ("fluid_euler3", "euler"), ("gas_pc3", "pc")])
def test_time_steps_match_golden(name, integrator, lut):
    base = name.split("_")[0]
    i, o = golden(f"{base}_in.snap"), golden(f"{name}.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    eng = Engine(setup, len(i["mass"]))
    eng.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
    eng.set_last_timestep(consts["initial_dt"])
    dts = o["dt_history"]
    for s in range(len(dts) - 1):
        if integrator == "pc":
            dt, _, _ = eng.step_pc(float(dts[s]), consts["max_dt"])
        else:
            eng.integrate()
            eng.euler(float(dts[s]))
            dt, _ = eng.compute_timestep(consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1], (s, dt, dts[s + 1])
    got = eng.download_state([k for k in ("pos", "vel", "rho", "u", "S", "damage", "acc", "du", "drho", "dS") if k in o])
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)
    eng.close()


def test_ghost_particles_are_neighbours_only(lut):
    """Owned particles + appended ghosts (halo) reproduce the single-domain derivatives of the owned ones."""
    i, o = golden("hello_in.snap"), golden("hello_out.snap")
    n = len(i["mass"])
    owned = np.where(i["pos"][:, 0] < 0)[0]
    ghost = np.where(i["pos"][:, 0] >= 0)[0]
    perm = np.concatenate([owned, ghost])
    part = {k: (v[perm] if (hasattr(v, "shape") and v.shape[:1] == (n,)) else v) for k, v in i.items()}
    setup = abi.setup_from_snapshot(i, lut)
    setup.materials[0].begin, setup.materials[0].end = 0, len(owned)
    eng = Engine(setup, len(owned), capacity=n)
    own = {k: v[: len(owned)] for k, v in part.items() if k in STATE_IN}
    gh = {k: v[len(owned):] for k, v in part.items() if k in STATE_IN}
    eng.upload_state(own, STATE_IN)
    eng.upload_state(gh, STATE_IN, first=len(owned))
    eng.upload("MATERIAL_ID", 0, np.zeros(len(ghost), np.uint32), first=len(owned))
    eng.set_active(n)
    eng.integrate()
    got = eng.download_state(["acc", "du", "drho", "dS", "divv", "ncnt"])
    assert np.array_equal(got["ncnt"], o["ncnt"][owned])
    for k in ("acc", "du", "drho", "dS", "divv"):
        assert_close(k, got[k], o[k][owned], TOL, FLOOR)
    eng.close()


def test_variants_agree_on_seeded_random_input(lut):
    """Direct and tiled kernels on a seeded random cloud (ragged cells, strongly varying h) against the oracle."""
    rng = np.random.default_rng(1234)
    i = golden("collision_in.snap")
    n = len(i["mass"])
    snap = dict(i)
    snap["pos"] = i["pos"].copy()
    snap["pos"][:, :3] += rng.normal(0, 0.4, (n, 3)) * i["pos"][:, 3:4]
    snap["pos"][:, 3] *= rng.uniform(0.6, 1.8, n)
    setup = abi.setup_from_snapshot(i, lut)
    orc = OraclePort(snap, setup)
    orc.integrate()
    for variant in (0, 1, 2, 3):
        eng, stats = gpu_integrate(snap, setup, variant)
        check_against(eng, stats, orc.a)
        eng.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args", [
    ["--config", "hello", "--n", 10000, "--solver", "asym"],                       # BASELINE configs[0]
    ["--config", "collision_preset", "--n", 100000, "--jitter", 21],               # configs[1] scale, jittered
    ["--config", "collision", "--n", 100000, "--solver", "asym"],                  # configs[1], examples/04 settings
    ["--config", "preset", "--n", 200000],                                         # configs[2] family
    ["--config", "preset_const_h", "--n", 30000, "--jitter", 5],
    ["--config", "fluid", "--n", 100000, "--jitter", 22],                          # configs[4] family
    ["--config", "collision_preset", "--n", 20000, "--jitter", 14, "--enforcing"],            # SOUND_SPEED_ENFORCING
    ["--config", "collision_preset", "--n", 20000, "--jitter", 15, "--continuity-undamaged"], # ContinuityEnum mode
    ["--config", "collision_preset", "--n", 20000, "--jitter", 16, "--sum-all", "--corrected", 0],
    ["--config", "hello", "--n", 20000, "--solver", "asym", "--jitter", 17, "--const-h", "--corrected", 1],
])
def test_against_live_reference(args, tmp_path):
    i, o = run_ref(str(tmp_path), args + ["--neighbours"])
    eng, stats = gpu_integrate(i, abi.setup_from_snapshot(i))
    check_against(eng, stats, o)
    off, idx = eng.neighbours()
    assert np.array_equal(off, o["nbr_offsets"]) and np.array_equal(idx, o["nbr_idx"])
    eng.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_predictor_corrector_against_live_reference(tmp_path):
    args = ["--config", "collision_preset", "--n", 20000, "--steps", 5]
    i, o = run_ref(str(tmp_path), args)
    setup = abi.setup_from_snapshot(i)
    consts = abi.run_constants(i)
    eng = Engine(setup, len(i["mass"]))
    eng.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
    eng.set_last_timestep(consts["initial_dt"])
    dts = o["dt_history"]
    for s in range(len(dts) - 1):
        dt, _, _ = eng.step_pc(float(dts[s]), consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1]
    got = eng.download_state(["pos", "vel", "rho", "u", "S", "damage"])
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)
    eng.close()


def test_empty_and_tiny_inputs(lut):
    """Edge cases the reference's finder tests cover (finders/test/Finders.cpp:23-54): empty storage, single particle."""
    i = golden("fluid_in.snap")
    for n in (0, 1, 2):
        setup = abi.setup_from_snapshot(i, lut)
        setup.materials[0].begin, setup.materials[0].end = 0, n
        eng = Engine(setup, n)
        eng.upload_state({k: v[:n] for k, v in i.items() if k in STATE_IN}, STATE_IN)
        if n == 0:
            eng.upload("MASS", 0, np.zeros(0))
        st = eng.integrate()
        assert st.pair_count == 0 or n == 2
        eng.close()


def test_cpp_dropin_through_isolver():
    """The C++ drop-in (opensph_b200/host/GpuSolver.cpp, built against the reference headers) next to the reference's
    own AsymmetricSolver / PredictorCorrector on identical Storages; see tests/dropin/dropin_test.cpp."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference at build time)")
    out = subprocess.run([exe, "20000", "3"], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and "DROPIN PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-500:]


def _oracle_vs_gpu(snap, setup, names=("acc", "du", "drho", "dS", "divv", "gradv")):
    orc = OraclePort(snap, setup)
    orc.integrate()
    eng, stats = gpu_integrate(snap, setup)
    got = eng.download_state([k for k in names if k in orc.a] + ["ncnt"])
    eng.close()
    assert np.array_equal(got["ncnt"], orc.a["ncnt"])
    for k in names:
        if k in orc.a:
            assert_close(k, got[k], orc.a[k], TOL, FLOOR)
    return got, orc


def test_huge_coordinates(lut):
    """Finder robustness test of the reference (finders/test/Finders.cpp: 'huge coordinates'): the same cloud far from
    the origin must give the same neighbour sets (the FP32 pre-filter works on unit-local coordinates)."""
    i = golden("collision_in.snap")
    setup = abi.setup_from_snapshot(i, lut)
    snap = dict(i)
    snap["pos"] = i["pos"].copy()
    snap["pos"][:, :3] += np.array([3.0e9, -7.0e9, 1.0e10])
    got, orc = _oracle_vs_gpu(snap, setup)
    base = OraclePort(i, setup)
    base.integrate()
    # at 1e10 m the coordinates only resolve ~1e-6 m, so a pair within that of the cut-off may flip; counts must agree
    # between GPU and oracle on the SAME shifted input (asserted above) and be close to the unshifted ones
    assert np.abs(got["ncnt"].astype(int) - base.a["ncnt"].astype(int)).max() <= 1


def test_line_of_particles_with_increasing_h(lut):
    """Finders.cpp 'increasing-h line': particles on a line, h growing with x, so search radii differ strongly; also a
    degenerate grid (one cell row) for the tiled kernel."""
    i = golden("fluid_in.snap")
    n = 300
    setup = abi.setup_from_snapshot(i, lut)
    setup.materials[0].begin, setup.materials[0].end = 0, n
    x = np.cumsum(np.linspace(1.0, 6.0, n))
    snap = {k: v[:n].copy() for k, v in i.items() if hasattr(v, "shape") and v.shape[:1] == (len(i["mass"]),)}
    snap["pos"][:, 0], snap["pos"][:, 1], snap["pos"][:, 2] = x, 0.0, 0.0
    snap["pos"][:, 3] = np.linspace(1.5, 9.0, n)
    snap["vel"][:, :3] = np.stack([np.sin(x / 50.0), np.zeros(n), np.zeros(n)], axis=1) * 10.0
    _oracle_vs_gpu(snap, setup, names=("acc", "du", "drho", "divv"))


@pytest.mark.parametrize("variant", [0, 2, 3])
@pytest.mark.parametrize("giants", [1, 1500])
def test_one_giant_smoothing_length(variant, giants, lut):
    """Particles whose h covers the whole cloud; each has every other particle as a neighbour.
    giants = 1: the two-level search radii (the device's answer to AsymmetricSolver's RadiiHashMap,
    AsymmetricSolver.cpp:14-56) keep it out of the cell list, the cell edge follows the ordinary particles.
    giants = 1500 (> LARGE_MAX): no split; the grid degenerates to a few cells with thousands of particles each (cell rows
    are then no longer ordered by x and are scanned whole, chunks are split into pieces, candidate lists overflow many
    times: several list blocks per chunk)."""
    i = golden("fluid_in.snap")
    n0, reps = len(i["mass"]), 12
    n = n0 * reps
    rng = np.random.default_rng(77)
    setup = abi.setup_from_snapshot(i, lut)
    setup.materials[0].begin, setup.materials[0].end = 0, n
    snap = {k: np.concatenate([v] * reps) for k, v in i.items() if hasattr(v, "shape") and v.shape[:1] == (n0,)}
    ext = np.ptp(i["pos"][:, :3], axis=0).max()
    shift = rng.uniform(0, 2.0 * ext, (reps, 3))
    snap["pos"] = snap["pos"].copy()
    snap["pos"][:, :3] += np.repeat(shift, n0, axis=0) + rng.normal(0, 0.05, (n, 3)) * snap["pos"][:, 3:4]
    snap["pos"][rng.choice(n, giants, replace=False), 3] = 4.0 * ext
    snap["vel"] = snap["vel"].copy()
    snap["vel"][:, :3] = np.sin(snap["pos"][:, :3] / ext * 3.0) * 5.0
    orc = OraclePort(snap, setup)
    orc.integrate()
    assert orc.a["ncnt"].max() == n - 1
    eng, stats = gpu_integrate(snap, setup, variant)
    got = eng.download_state(["acc", "du", "drho", "divv", "ncnt"])
    off, idx = eng.neighbours()
    ooff, oidx = orc.neighbours()
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx)
    eng.close()
    assert np.array_equal(got["ncnt"], orc.a["ncnt"])
    for k in ("acc", "du", "drho", "divv"):
        assert_close(k, got[k], orc.a[k], TOL, FLOOR)


def test_giant_smoothing_length_costs_little_at_bench_scale():
    """one huge particle in the 1 M-particle lattice must not set the cell edge for everybody. The step time
    with the giant stays within 2x of the uniform-h time (it was ~100x with a single-level grid), the neighbour count of
    every ordinary particle rises by exactly one and the giant sees all others."""
    from opensph_b200 import workloads
    state = workloads.basalt_sphere_state(1_000_000)
    n = len(state["mass"])
    setup = workloads.make_setup(n)
    names = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
             "eps_min", "m_zero", "growth", "n_flaws", "flag")
    times, ncnt = {}, {}
    for giant in (False, True):
        st = {k: v.copy() for k, v in state.items()}
        if giant:
            st["pos"][n // 2, 3] = 4.0 * 5.0e4
        eng = Engine(setup, n)
        eng.upload_state(st, names)
        eng.integrate()
        best = min(eng.integrate().gpu_ms for _ in range(3))
        times[giant] = best
        ncnt[giant] = eng.download_state(["ncnt"])["ncnt"]
        eng.close()
    assert ncnt[True][n // 2] == n - 1
    # every other particle gains the giant as a neighbour -- except those that had particle n/2 as a neighbour already
    gain = ncnt[True].astype(np.int64) - ncnt[False].astype(np.int64)
    gain[n // 2] = 1
    assert set(np.unique(gain)) <= {0, 1}
    assert int((gain == 0).sum()) == int(ncnt[False][n // 2])
    assert times[True] <= 2.0 * times[False], times


def test_conservation_at_full_bench_size():
    """Size-independent property at BASELINE's headline size (10.6 M particles, the bench workload after a few moving
    PredictorCorrector steps): the pair forces are antisymmetric (symmetrised h, p_i/rho_i^2 + p_j/rho_j^2 + Pi_ij, the
    undamaged-pair filter is symmetric), so sum_i m_i a_i vanishes to rounding; the neighbour relation is symmetric, so the
    sum of the neighbour counts is even and equals the pair count of the statistics. A candidate dropped by the u16 list
    indices, the 32-bit cell ranges or the list pool at this size would break either."""
    from opensph_b200 import workloads
    state = workloads.basalt_sphere_state(10_000_000)
    n = len(state["mass"])
    assert n > 10_000_000
    setup = workloads.make_setup(n)
    names = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
             "eps_min", "m_zero", "growth", "n_flaws", "flag")
    eng = Engine(setup, n)
    eng.upload_state(state, names)
    _, _, st = eng.run_pc(4, 0.01, 10.0)  # the lists of step 1 are reused with moved particles
    got = eng.download_state(["acc", "ncnt"])
    eng.close()
    total = int(got["ncnt"].astype(np.int64).sum())
    assert total == st.pair_count and total % 2 == 0
    assert 60.0 < total / n < 70.0
    force = state["mass"][:, None] * got["acc"][:, :3]
    residual = np.abs(force.sum(axis=0)) / np.abs(force).sum(axis=0)
    assert np.all(residual < 1e-12), residual


def test_batched_steps_equal_single_steps(lut):
    """sphgpu_run_pc (time step chosen on the device, one host synchronisation for the batch) against the same number
    of sphgpu_step_pc calls with the host feeding the time step back: identical dt sequence and state."""
    i = golden("collision_in.snap")
    setup = abi.setup_from_snapshot(i, lut)
    consts = abi.run_constants(i)
    n, steps = len(i["mass"]), 6
    a, b = Engine(setup, n), Engine(setup, n)
    for eng in (a, b):
        eng.upload_state(i, STATE_IN)
        eng.set_last_timestep(consts["initial_dt"])
    dts, dt = [], consts["initial_dt"]
    for _ in range(steps):
        dt, crit, _ = a.step_pc(dt, consts["max_dt"])
        dts.append(dt)
    hist, crits, st = b.run_pc(steps, consts["initial_dt"], consts["max_dt"])
    assert np.array_equal(hist, np.array(dts)), (hist, dts)
    ga, gb = (e.download_state(["pos", "vel", "rho", "u", "S", "damage", "acc", "du", "drho", "dS"]) for e in (a, b))
    for k in ga:
        assert np.array_equal(ga[k], gb[k]), k
    # and the batch can be continued by single steps
    dt_a, _, _ = a.step_pc(dts[-1], consts["max_dt"])
    dt_b, _, _ = b.step_pc(float(hist[-1]), consts["max_dt"])
    assert dt_a == dt_b
    a.close()
    b.close()


def test_list_pool_overflow_falls_back_and_grows(lut):
    """Thousands of neighbours per particle (h three times the spacing-consistent value): the candidate lists need far
    more rows than the pool reserves, so the first evaluation takes the fused path for the overflowing units
    (stats.reserved0 > 0) and enlarges the pool; results must be right both times."""
    i = golden("fluid_in.snap")
    n0, reps = len(i["mass"]), 6
    n = n0 * reps
    rng = np.random.default_rng(5)
    setup = abi.setup_from_snapshot(i, lut)
    setup.materials[0].begin, setup.materials[0].end = 0, n
    snap = {k: np.concatenate([v] * reps) for k, v in i.items() if hasattr(v, "shape") and v.shape[:1] == (n0,)}
    snap["pos"] = snap["pos"].copy()
    snap["pos"][:, :3] += rng.normal(0, 0.3, (n, 3)) * snap["pos"][:, 3:4]  # the copies overlap: six times the density
    snap["pos"][:, 3] *= 3.0
    orc = OraclePort(snap, setup)
    orc.integrate()
    assert orc.a["ncnt"].mean() > 1000
    eng = Engine(setup, n)
    eng.upload_state(snap, STATE_IN)
    first = eng.integrate()
    assert first.reserved0 > 0, "the pool was expected to overflow"
    for attempt in range(4):
        got = eng.download_state(["acc", "du", "drho", "divv", "ncnt"])
        assert np.array_equal(got["ncnt"], orc.a["ncnt"])
        for k in ("acc", "du", "drho", "divv"):
            assert_close(k, got[k], orc.a[k], TOL, FLOOR)
        again = eng.integrate()
        if again.reserved0 == 0:
            break
    assert again.reserved0 < first.reserved0, "the pool did not grow"
    eng.close()


def test_constant_velocity_gives_exactly_zero_gradient(lut):
    """equations/test/EquationTerm.cpp:177-375: the gradient of a constant velocity field is EXACTLY zero."""
    i = golden("hello_in.snap")
    setup = abi.setup_from_snapshot(i, lut)
    snap = dict(i)
    snap["vel"] = i["vel"].copy()
    snap["vel"][:, :3] = np.array([12.5, -3.0, 7.25])
    eng, _ = gpu_integrate(snap, setup)
    got = eng.download_state(["divv", "gradv"])
    eng.close()
    assert np.all(got["divv"] == 0.0) and np.all(got["gradv"] == 0.0)


def test_symmetric_solver_golden_on_gpu(lut):
    """the reference's SymmetricSolver output (golden hello_sym_out.snap, made with --solver sym) equals the
    asymmetric evaluation (the reference's own cross-check, solvers/test/Solvers.cpp:178-216); the GPU path must reproduce
    it within the parity tolerance as well."""
    i, o = golden("hello_in.snap"), golden("hello_sym_out.snap")
    eng, stats = gpu_integrate(i, abi.setup_from_snapshot(i, lut))
    got = eng.download_state(["acc", "du", "drho", "dS", "divv", "ncnt"])
    if "ncnt" in o:
        assert np.array_equal(got["ncnt"], o["ncnt"])
    for k in ("acc", "du", "drho", "dS", "divv"):
        assert_close(k, got[k], o[k], TOL, FLOOR)
    eng.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_bench_scale_preset_against_live_reference(tmp_path):
    """BASELINE configs[2] at its full size: the 1 M-particle collision-preset lattice (jittered so that lattice sums do
    not cancel to zero), one integrate() of the unmodified reference against the GPU path: NEIGHBOR_CNT exactly, the
    derivatives within tolerance. Exercises the code paths small inputs never reach (thousands of work units, list pool
    near its design load, 32-bit cell ranges)."""
    i, o = run_ref(str(tmp_path), ["--config", "preset", "--n", 1000000, "--jitter", 31])
    assert len(i["mass"]) > 1_000_000
    eng, stats = gpu_integrate(i, abi.setup_from_snapshot(i))
    # a million particles sample the cancellation of (1 - D^3) in VonMisesRheology::initialize ten times closer to D = 1
    # than the 100 k cases do: the stress-reducing factor agrees to 2e-12 instead of 1e-12
    check_against(eng, stats, o, tol_exact=1e-11)
    eng.close()


def test_list_reuse_matches_rebuild_every_step(lut):
    """The candidate lists are conservative supersets built with a skin; the exact predicate runs every step. A run that
    reuses them (device-side displacement check decides when to rebuild) must give the same neighbour counts and, up to
    summation order, the same state as a run that rebuilds in every step -- while the particles really move."""
    from opensph_b200 import workloads
    state = workloads.basalt_sphere_state(20000)
    state["vel"] = state["vel"] * 60.0  # ~3 km/s: a step moves particles by a few percent of h
    n = len(state["mass"])
    setup = workloads.make_setup(n)
    names = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
             "eps_min", "m_zero", "growth", "n_flaws", "flag")
    out = {}
    for skin in (0.0, 0.04):
        eng = Engine(setup, n)
        eng.set_list_skin(skin)
        eng.upload_state(state, names)
        steps = 40
        ncnts = []
        dts, _, _ = eng.run_pc(steps // 2, 0.01, 10.0)
        ncnts.append(eng.download_state(["ncnt"])["ncnt"])
        dts2, _, _ = eng.run_pc(steps // 2, float(dts[-1]), 10.0)
        ncnts.append(eng.download_state(["ncnt"])["ncnt"])
        out[skin] = (eng.download_state(["pos", "vel", "rho", "u", "S", "damage", "acc", "du", "drho", "dS"]), np.concatenate([dts, dts2]),
                     eng.list_stats(), ncnts)
        eng.close()
    (a, dta, sa, na), (b, dtb, sb, nb) = out[0.0], out[0.04]
    moved = np.abs(a["pos"][:, :3] - state["pos"][:, :3]).max() / state["pos"][0, 3]
    assert moved > 0.1, f"the test must move the particles (max displacement {moved:.3f} h)"
    assert sa[0] == 40, sa
    assert 2 <= sb[0] < 30, f"expected a few rebuilds with the skin, got {sb}"
    for x, y in zip(na, nb):
        assert np.array_equal(x, y), "NEIGHBOR_CNT differs between reused and rebuilt lists"
    assert np.allclose(dta, dtb, rtol=1e-11, atol=0)
    for k in a:
        assert_close(k, b[k], a[k], 1e-9, FLOOR)


def test_list_reuse_relative_displacement_criterion(lut):
    """The rebuild decision looks at the RELATIVE displacement of particles that can reach each other (per-cell displacement
    boxes widened over the candidate stencil plus one cell, api: sphgpu_set_list_skin). Two checks:
    (1) two bodies flying into each other -- neighbour sets between them change from step to step: NEIGHBOR_CNT after
        every step equals the run that rebuilds every step (no neighbour may be missed);
    (2) one body that translates fast and uniformly: no relative motion beyond the physical one, so the lists survive
        far longer than the absolute displacement (many h) would have allowed."""
    from opensph_b200 import workloads
    names = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
             "eps_min", "m_zero", "growth", "n_flaws", "flag")
    # (1) a sphere cut in two halves that approach each other at 8 km/s along x, with a little shear
    state = workloads.basalt_sphere_state(12000)
    h = state["pos"][0, 3]
    left = state["pos"][:, 0] < 0.0
    state["pos"][left, 0] -= 0.8 * h
    state["pos"][~left, 0] += 0.8 * h
    state["vel"][:, :3] = 0.0
    state["vel"][left, 0], state["vel"][~left, 0] = 4000.0, -4000.0
    state["vel"][left, 1] = 500.0
    n = len(state["mass"])
    setup = workloads.make_setup(n)
    counts = {}
    for skin in (0.0, 0.03):
        with Engine(setup, n) as eng:
            eng.set_list_skin(skin)
            eng.upload_state(state, names)
            dt, seq = 0.05, []
            for _ in range(30):
                dts, _, _ = eng.run_pc(1, dt, 0.1)
                dt = float(dts[-1])
                seq.append(eng.download_state(["ncnt"])["ncnt"].copy())
            counts[skin] = (seq, eng.list_stats())
    for k, (x, y) in enumerate(zip(counts[0.0][0], counts[0.03][0])):
        assert np.array_equal(x, y), f"NEIGHBOR_CNT differs at step {k} between reused and rebuilt lists"
    assert counts[0.0][0][-1].sum() != counts[0.0][0][0].sum(), "the halves must have come into contact"
    assert counts[0.03][1][0] < 30, counts[0.03][1]
    # (2) uniform translation at 30 km/s: after 20 steps of 50 ms every particle has moved ~6 h
    state = workloads.basalt_sphere_state(12000)
    state["vel"][:, :3] *= 0.02
    state["vel"][:, 0] += 30000.0
    n = len(state["mass"])
    setup = workloads.make_setup(n)
    with Engine(setup, n) as eng:
        eng.set_list_skin(0.03)
        eng.upload_state(state, names)
        eng.run_pc(20, 0.05, 0.05)
        moved = np.abs(eng.download_state(["pos"])["pos"][:, 0] - state["pos"][:, 0]).min() / state["pos"][0, 3]
        rebuilds = eng.list_stats()[0]
        a = eng.download_state(["ncnt", "acc"])
    with Engine(setup, n) as eng:
        eng.set_list_skin(0.0)
        eng.upload_state(state, names)
        eng.run_pc(20, 0.05, 0.05)
        b = eng.download_state(["ncnt", "acc"])
    assert moved > 3.0, moved
    assert rebuilds <= 3, f"a uniformly translating body must keep its lists ({rebuilds} builds in 20 steps)"
    assert np.array_equal(a["ncnt"], b["ncnt"])
    assert_close("acc", a["acc"], b["acc"], 1e-9, FLOOR)


def test_multi_gpu_parity_script():
    """tests/run_mgpu_parity.py (two ranks: NCCL halo exchange, batched stepping, repartition) against a single-domain run;
    needs two GPUs on the box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from conftest import ROOT
    env = dict(os.environ, MGPU_PARTICLES="120000")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "tests", "run_mgpu_parity.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU PARITY OK" in r.stdout, r.stdout[-3000:]


def test_async_transfers_equal_synchronous_ones(lut):
    """sphgpu_upload_async / sphgpu_download_async (queued copies, downloads on a second stream) against the synchronous
    calls: identical bytes, also when several batches are in flight before one sphgpu_transfer_sync."""
    i = golden("collision_in.snap")
    setup = abi.setup_from_snapshot(i, lut)
    n = len(i["mass"])
    a, b = Engine(setup, n), Engine(setup, n)
    a.upload_state(i, STATE_IN)
    for k in STATE_IN:
        q, order = abi.SNAPSHOT_FIELDS[k]
        b.upload_async(q, order, np.ascontiguousarray(i[k], dtype=abi.QUANTITIES[q][2]))
    sa, sb = a.integrate(), b.integrate()
    assert sa.pair_count == sb.pair_count
    names = ["acc", "du", "drho", "dS", "divv", "ncnt", "pos"]
    ref = a.download_state(names)
    outs = []
    for batch in range(3):  # three batches queued back to back
        out = {k: np.empty_like(ref[k]) for k in names}
        for k in names:
            q, order = abi.SNAPSHOT_FIELDS[k]
            b.download_async(q, order, out[k])
        b.download_batch_end()
        outs.append(out)
    b.transfer_sync()
    for out in outs:
        for k in names:
            assert np.array_equal(out[k], ref[k]), k
    a.close()
    b.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("config", ["collision_preset", "fluid"])
def test_balsara_switch_against_live_reference(config, tmp_path):
    """BalsaraSwitch<StandardAV> (core/sph/equations/av/Balsara.h). One evaluation (the factor is built
    from the PREVIOUS evaluation's div v and rot v, so it is zero here and the viscosity is off) and five
    PredictorCorrector steps, in which the factors are live, against the reference's own integrator."""
    i, o = run_ref(str(tmp_path), ["--config", config, "--n", 20000, "--jitter", 18, "--balsara"])
    setup = abi.setup_from_snapshot(i)
    assert setup.cfg.flags & abi.FLAG_BALSARA
    eng, stats = gpu_integrate(i, setup)
    check_against(eng, stats, o)
    eng.close()
    i, o = run_ref(str(tmp_path), ["--config", config, "--n", 20000, "--steps", 5, "--balsara"])
    setup = abi.setup_from_snapshot(i)
    consts = abi.run_constants(i)
    eng = Engine(setup, len(i["mass"]))
    eng.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
    eng.set_last_timestep(consts["initial_dt"])
    dts = o["dt_history"]
    for s in range(len(dts) - 1):
        dt, _, _ = eng.step_pc(float(dts[s]), consts["max_dt"])
        assert abs(dt - dts[s + 1]) <= 1e-9 * dts[s + 1]
    got = eng.download_state([k for k in ("pos", "vel", "rho", "u", "S", "damage") if k in o])
    for k, v in got.items():
        assert_close(k, v, o[k], 1e-9, FLOOR)
    # the switch must matter in this comparison: without it the state differs
    plain = abi.setup_from_snapshot(i)
    plain.cfg.flags &= ~abi.FLAG_BALSARA
    ref = Engine(plain, len(i["mass"]))
    ref.upload_state(i, STATE_IN + ("acc", "drho", "du", "dS", "ddamage"))
    ref.set_last_timestep(consts["initial_dt"])
    for s in range(len(dts) - 1):
        ref.step_pc(float(dts[s]), consts["max_dt"])
    other = ref.download_state(["vel"])["vel"]
    assert np.abs(other[:, :3] - got["vel"][:, :3]).max() > 1e-6 * np.abs(got["vel"][:, :3]).max()
    eng.close()
    ref.close()
