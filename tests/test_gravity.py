"""Self-gravity: the plain-C oracle and the product's host-visible arithmetic against golden vectors of
the reference's BruteForceGravity / BarnesHut (CPU tests), and the device Barnes-Hut through the C ABI (-m gpu).

What can be bit-close and what cannot: with every node opened (opening angle <= 0) the device evaluates exactly the pair
sums of BruteForceGravity, so it must match to rounding (1e-10). With an opening angle the device's tree (a binary radix
tree over Morton-sorted particles) differs from the reference's k-d tree, so the two Barnes-Hut results agree only within
the error of the multipole approximation -- the reference's own tests measure BarnesHut against BruteForceGravity the
same way (core/gravity/test/BarnesHut.cpp). The bar used here: the device's error against brute force must not exceed
twice the (RMS) error the reference's BarnesHut makes on the same input with the same settings; the error of the worst
particle, which depends on where the tree happens to cut, four times."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import REF_FAST, REF_STRICT, ROOT, golden, have_ref
from compare import assert_close
from opensph_b200 import abi
from opensph_b200.snapshot import read_snapshot
import oracle_port as op

HOST_SRC = os.path.join(ROOT, "tests", "csrc", "host_math_check.cpp")
HOST_LIB = os.path.join(ROOT, "tests", "csrc", "libhostcheck.so")
_D = C.POINTER(C.c_double)


def rms_err(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def max_err(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def hostcheck():
    deps = [HOST_SRC, os.path.join(ROOT, "opensph_b200", "csrc", "sph_math.cuh"), os.path.join(ROOT, "opensph_b200", "csrc", "grav_math.cuh")]
    if not os.path.exists(HOST_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOST_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", HOST_SRC, "-o", HOST_LIB])
    return C.CDLL(HOST_LIB)


# ---- CPU: the oracle against the reference's golden vectors ------------------------------------------------------------
def test_oracle_gravity_table_matches_reference():
    g = golden("gravity_brute.snap")
    assert np.abs(op.gravity_lut(40000) - g["grav_lut"]).max() <= 4e-15
    assert np.abs(abi.gravity_table_cubic_spline(40000) - g["grav_lut"]).max() <= 4e-15  # what bench / workloads use


@pytest.mark.parametrize("name,point", [("gravity_brute", False), ("gravity_brute_point", True)])
def test_oracle_brute_force_matches_reference(name, point):
    g = golden(f"{name}.snap")
    lut = None if point else golden("gravity_brute.snap")["grav_lut"]
    acc = op.gravity_brute(g["pos"], g["mass"], g["grav_params"][2], lut, 0.0 if point else g["grav_params"][3])
    assert max_err(acc, g["grav_acc"][:, :3]) <= 1e-13


def test_oracle_multipoles_match_reference():
    g = golden("gravity_bh.snap")
    com, mom = op.gravity_moments(g["pos"], g["mass"])
    ref = g["root_moments"]
    # the reference carries its traceless tensors through the k-d tree with the parallel-axis theorem: rounding only
    scale2, scale3 = np.abs(ref[1:6]).max(), np.abs(ref[6:13]).max()
    assert abs(mom[0] - ref[0]) <= 1e-13 * ref[0]
    assert np.abs(mom[1:6] - ref[1:6]).max() <= 1e-10 * scale2
    assert np.abs(mom[6:13] - ref[6:13]).max() <= 1e-10 * scale3
    G = g["grav_params"][2]
    for pp, pa in zip(g["probe_pos"], g["probe_acc"]):
        acc = op.gravity_multipole(com, mom * G, int(pp[3]), pp)
        assert np.abs(acc - pa).max() <= 1e-11 * np.abs(pa).max()


def test_reference_barnes_hut_error_is_what_the_bar_assumes():
    # documents the size of the approximation error the GPU tests are measured against
    bh, br = golden("gravity_bh.snap"), golden("gravity_brute.snap")
    assert 1e-6 < rms_err(bh["grav_acc"][:, :3], br["grav_acc"][:, :3]) < 2e-3


# ---- CPU: the product's arithmetic (grav_math.cuh) against the oracle ---------------------------------------------------
@pytest.mark.parametrize("pieces", [1, 7, 64])
@pytest.mark.parametrize("order", [0, 2, 3])
def test_product_moment_math_matches_oracle(hostcheck, pieces, order):
    g = golden("gravity_bh.snap")
    pos, mass = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["mass"])
    G = g["grav_params"][2]
    com_o, mom_o = op.gravity_moments(pos, mass)
    probes = np.ascontiguousarray(g["probe_pos"][:4, :3])
    com, mom, acc = np.zeros(3), np.zeros(13), np.zeros((4, 3))
    gm = np.ascontiguousarray(mass * G)
    hostcheck.hostcheck_gravity_moments(C.c_uint32(len(mass)), pos.ctypes.data_as(_D), gm.ctypes.data_as(_D), C.c_uint32(pieces),
                                        com.ctypes.data_as(_D), mom.ctypes.data_as(_D), C.c_uint32(4), probes.ctypes.data_as(_D),
                                        C.c_int(order), acc.ctypes.data_as(_D))
    assert np.abs(com - com_o).max() <= 1e-12 * np.abs(pos[:, :3]).max()
    assert abs(mom[0] - G * mom_o[0]) <= 1e-13 * G * mom_o[0]
    assert np.abs(mom[1:6] - G * mom_o[1:6]).max() <= 1e-10 * G * np.abs(mom_o[1:6]).max()
    assert np.abs(mom[6:13] - G * mom_o[6:13]).max() <= 1e-10 * G * np.abs(mom_o[6:13]).max()
    for k in range(4):
        want = op.gravity_multipole(com_o, mom_o * G, order, probes[k])
        assert np.abs(acc[k] - want).max() <= 1e-11 * np.abs(want).max()


@pytest.mark.parametrize("name,point", [("gravity_brute", False), ("gravity_brute_point", True)])
def test_product_pair_math_matches_reference(hostcheck, name, point):
    g = golden(f"{name}.snap")
    pos, mass = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["mass"])
    lut = np.ascontiguousarray(golden("gravity_brute.snap")["grav_lut"])
    acc = np.zeros((len(mass), 3))
    hostcheck.hostcheck_gravity_pairs(C.c_uint32(len(mass)), pos.ctypes.data_as(_D), mass.ctypes.data_as(_D), C.c_double(g["grav_params"][2]),
                                      lut.ctypes.data_as(_D), C.c_uint32(len(lut) - 1), C.c_double(0.0 if point else 2.0),
                                      acc.ctypes.data_as(_D))
    assert max_err(acc, g["grav_acc"][:, :3]) <= 1e-12


# ---- GPU ------------------------------------------------------------------------------------------------------------------
def _engine_with_positions(pos, mass, lut):
    """An engine whose only meaningful state is {r, h, m}: the fluid golden setup serves as the container."""
    from opensph_b200.engine import Engine
    snap = golden("fluid_in.snap")
    setup = abi.setup_from_snapshot(snap, lut)
    n = len(mass)
    setup.materials[0].begin, setup.materials[0].end = 0, n
    eng = Engine(setup, n, device=0)
    eng.upload("POSITION", 0, pos)
    eng.upload("MASS", 0, mass)
    return eng


def _gpu_gravity(pos, mass, lut, theta, order, G, point=False, leaf=0, grav_lut=None):
    with _engine_with_positions(pos, mass, lut) as eng:
        eng.gravity_configure(theta, order, G, None if point else grav_lut, 0.0 if point else 2.0, leaf)
        st = eng.gravity_eval(accumulate=False)
        acc = eng.download("POSITION", 2)[:, :3].copy()
    return acc, st


@pytest.mark.gpu
@pytest.mark.parametrize("name,point", [("gravity_brute", False), ("gravity_brute_point", True)])
def test_gpu_exact_mode_matches_brute_force_golden(lut, name, point):
    g = golden(f"{name}.snap")
    grav_lut = golden("gravity_brute.snap")["grav_lut"]
    acc, st = _gpu_gravity(g["pos"], g["mass"], lut, 0.0, 3, g["grav_params"][2], point, 0, grav_lut)
    assert st.approximated == 0 and st.exact > 0
    assert_close("gravity", acc, g["grav_acc"][:, :3], 1e-10, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name,brute,point", [("gravity_bh", "gravity_brute", False), ("gravity_bh_point", "gravity_brute_point", True)])
def test_gpu_barnes_hut_within_the_reference_error(lut, name, brute, point):
    g, b = golden(f"{name}.snap"), golden(f"{brute}.snap")
    theta, order, G = g["grav_params"][0], int(g["grav_params"][1]), g["grav_params"][2]
    acc, st = _gpu_gravity(g["pos"], g["mass"], lut, theta, order, G, point, int(g["grav_params"][4]), golden("gravity_brute.snap")["grav_lut"])
    exact = b["grav_acc"][:, :3]
    ref_rms, ref_max = rms_err(g["grav_acc"][:, :3], exact), max_err(g["grav_acc"][:, :3], exact)
    assert st.approximated > 0
    assert rms_err(acc, exact) <= 2.0 * ref_rms, (rms_err(acc, exact), ref_rms)
    assert max_err(acc, exact) <= 4.0 * ref_max, (max_err(acc, exact), ref_max)  # the worst particle depends on the tree
    # and the two Barnes-Hut results agree with each other to the same order
    assert rms_err(acc, g["grav_acc"][:, :3]) <= 3.0 * ref_rms


@pytest.mark.gpu
def test_gpu_multipole_orders_converge(lut):
    # point particles: with the softening kernel the error of a node taken as a point mass inside the kernel's reach
    # (the reference accepts such nodes too) hides the truncation error of the expansion
    g, b = golden("gravity_bh_point.snap"), golden("gravity_brute_point.snap")
    G = g["grav_params"][2]
    errs = [rms_err(_gpu_gravity(g["pos"], g["mass"], lut, 0.7, o, G, True, 8)[0], b["grav_acc"][:, :3]) for o in (0, 2, 3)]
    assert errs[0] > 2.0 * errs[1] > 2.0 * errs[2] > 0.0, errs
    # a smaller opening angle is more accurate
    e_small = rms_err(_gpu_gravity(g["pos"], g["mass"], lut, 0.3, 3, G, True, 8)[0], b["grav_acc"][:, :3])
    assert e_small < 0.5 * errs[2]


@pytest.mark.gpu
def test_gpu_gravity_edge_cases(lut):
    grav_lut = golden("gravity_brute.snap")["grav_lut"]
    G = abi.GRAVITY_CONSTANT
    # one particle: no acceleration; two particles: Newton / the softened law
    for n in (1, 2, 3, 33):
        rng = np.random.default_rng(n)
        pos = np.concatenate([rng.uniform(-1.0, 1.0, (n, 3)), rng.uniform(0.05, 0.4, (n, 1))], axis=1)
        mass = rng.uniform(1.0, 2.0, n) * 1e9
        acc, _ = _gpu_gravity(pos, mass, lut, 0.5, 3, G, False, 0, grav_lut)
        want = op.gravity_brute(pos, mass, G, grav_lut, 2.0)
        if n == 1:
            assert np.all(acc == 0.0)
        else:
            assert np.abs(acc - want).max() <= 2e-3 * np.abs(want).max()
        accx, _ = _gpu_gravity(pos, mass, lut, 0.0, 3, G, False, 0, grav_lut)
        assert np.abs(accx - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-300)
    # coincident particles (identical Morton keys) and a wide dynamic range of scales
    rng = np.random.default_rng(5)
    cluster = rng.normal(0.0, 1e-3, (400, 3))
    far = rng.uniform(-1e3, 1e3, (200, 3))
    xyz = np.concatenate([cluster, far, np.zeros((40, 3)) + 7.0])
    pos = np.concatenate([xyz, np.full((len(xyz), 1), 1e-4)], axis=1)
    mass = np.full(len(xyz), 1e6)
    want = op.gravity_brute(pos, mass, G, grav_lut, 2.0)
    accx, _ = _gpu_gravity(pos, mass, lut, 0.0, 3, G, False, 0, grav_lut)
    assert np.abs(accx - want).max() <= 1e-10 * np.abs(want).max()
    acc, st = _gpu_gravity(pos, mass, lut, 0.5, 3, G, False, 0, grav_lut)
    assert st.approximated > 0
    assert rms_err(acc, want) <= 2e-3


@pytest.mark.gpu
def test_gpu_gravity_joins_the_sph_accelerations(lut):
    """sphgpu_integrate with gravity configured == SPH-only integrate + gravity (GravitySolver::loop, GravitySolver.cpp:64-99)."""
    from opensph_b200.engine import Engine
    snap = golden("collision_in.snap")
    setup = abi.setup_from_snapshot(snap, lut)
    names = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag")
    grav_lut = golden("gravity_brute.snap")["grav_lut"]
    with Engine(setup, len(snap["mass"]), device=0) as eng:
        eng.upload_state(snap, names)
        eng.integrate()
        sph = eng.download("POSITION", 2)[:, :3].copy()
        eng.gravity_configure(0.5, 3, abi.GRAVITY_CONSTANT * 1e6, grav_lut, 2.0)  # a constant large enough to matter next to the SPH terms
        eng.gravity_eval(accumulate=False)
        grav = eng.download("POSITION", 2)[:, :3].copy()
        eng.integrate()
        both = eng.download("POSITION", 2)[:, :3].copy()
        st = eng.gravity_last_stats()
        # three batched PredictorCorrector steps keep running with gravity
        hist = eng.run_pc(3, 1e-4, 1e-2)
        eng.gravity_off()
        eng.upload_state(snap, names)
        eng.integrate()
        again = eng.download("POSITION", 2)[:, :3].copy()
    assert st.groups > 0 and st.gpu_ms > 0.0
    assert np.abs(grav).max() > 1e-6 * np.abs(sph).max()
    assert np.abs(both - (sph + grav)).max() <= 1e-12 * np.abs(sph).max()
    assert np.array_equal(again, sph)
    assert len(hist[0]) == 3 and np.all(hist[0] > 0.0)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="needs the compiled reference (oracle/_ref)")
def test_gpu_barnes_hut_against_live_reference_100k(lut):
    """The collision-preset sphere at 100 k particles: device Barnes-Hut next to the reference's, both measured against the
    exact sums (device, opening angle 0); the times are printed for the record."""
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "g.snap")
        line = subprocess.check_output([REF_FAST, "gravity", "--config", "preset", "--n", "100000", "--jitter", "11", "--gravity", "bh", "--theta", "0.5",
                                        "--order", "3", "--leaf", "20", "--no-lut", "--out", out]).decode()
        g = read_snapshot(out)
    grav_lut = abi.gravity_table_cubic_spline(40000)
    G = g["grav_params"][2]
    exact, _ = _gpu_gravity(g["pos"], g["mass"], lut, 0.0, 3, G, False, 20, grav_lut)
    acc, st = _gpu_gravity(g["pos"], g["mass"], lut, 0.5, 3, G, False, 20, grav_lut)
    ref = g["grav_acc"][:, :3]
    e_ref, e_gpu = rms_err(ref, exact), rms_err(acc, exact)
    print(f"\n100k gravity: reference {line.strip()}  device {st.gpu_ms:.2f} ms, rms error vs exact: reference {e_ref:.2e}, device {e_gpu:.2e}, "
          f"node interactions {st.approximated}, exact ranges {st.exact}, groups {st.groups}")
    assert e_gpu <= 2.0 * e_ref
    assert max_err(acc, exact) <= 4.0 * max_err(ref, exact)
