"""compute-sanitizer driver (run on the GPU box):

    compute-sanitizer --tool memcheck  python tests/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py

Runs every tiled pair-kernel variant on the golden inputs (one work unit each), then a 20 k-particle lattice with
ghost-free multi-unit chains, and a cloud with one giant smoothing length (split chunks, multi-block lists)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import golden  # noqa: E402
from opensph_b200 import abi, workloads  # noqa: E402
from opensph_b200.engine import Engine  # noqa: E402

STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag")
lut = golden("lut.snap")


def run(snap, setup, variant, label):
    eng = Engine(setup, len(snap["mass"]))
    eng.set_variant(variant)
    eng.upload_state(snap, STATE_IN)
    st = eng.integrate()
    got = eng.download_state(["ncnt"])
    assert st.pair_count == int(got["ncnt"].astype(np.int64).sum())
    print(label, "variant", variant, "pairs", st.pair_count, "fallback units", st.reserved0)
    eng.close()


# a few PredictorCorrector steps with list reuse (cp.async prefetches, early-exit build kernels) and the Balsara switch
names = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
         "eps_min", "m_zero", "growth", "n_flaws", "flag")
small = workloads.basalt_sphere_state(6000, 5.0e4, solid=True)
small["vel"] = small["vel"] * 40.0


def section_pairs():
    for name in ("collision", "fluid"):
        i = golden(f"{name}_in.snap")
        for variant in (0, 2, 3):
            run(i, abi.setup_from_snapshot(i, lut), variant, name)

    state = workloads.basalt_sphere_state(20000, 5.0e4, solid=True)
    for variant in (0, 3):
        run(state, workloads.make_setup(len(state["mass"]), solid=True), variant, "lattice 20k")

    i = golden("fluid_in.snap")
    n0, reps = len(i["mass"]), 4
    n = n0 * reps
    rng = np.random.default_rng(77)
    setup = abi.setup_from_snapshot(i, lut)
    setup.materials[0].begin, setup.materials[0].end = 0, n
    snap = {k: np.concatenate([v] * reps) for k, v in i.items() if hasattr(v, "shape") and v.shape[:1] == (n0,)}
    ext = np.ptp(i["pos"][:, :3], axis=0).max()
    snap["pos"] = snap["pos"].copy()
    snap["pos"][:, :3] += np.repeat(rng.uniform(0, 2.0 * ext, (reps, 3)), n0, axis=0)
    snap["pos"][n // 2, 3] = 4.0 * ext
    run(snap, setup, 0, "giant h (two-level radii)")
    snap["pos"][rng.choice(n, 1200, replace=False), 3] = 4.0 * ext
    run(snap, setup, 0, "1200 giants (degenerate single-level grid)")



def section_steps():
    for balsara in (False, True):
        setup = workloads.make_setup(len(small["mass"]), solid=True)
        if balsara:
            setup.cfg.flags |= abi.FLAG_BALSARA
        eng = Engine(setup, len(small["mass"]))
        eng.set_list_skin(0.04)
        eng.upload_state(small, names)
        dts, _, st = eng.run_pc(8, 0.01, 10.0)
        print("run_pc 8 steps, balsara", balsara, "pairs", st.pair_count, "list builds / age / metric", eng.list_stats())
        eng.close()



def section_deltasph():
    # the delta-SPH records (176 / 144 bytes, two more staged pieces per neighbour): golden inputs in every tiled variant and a
    # few batched steps of a multi-unit lattice, solid and fluid
    for name in ("deltasph_in", "fluid_in"):
        i = golden(f"{name}.snap")
        for variant in (0, 2, 3):
            setup = abi.setup_from_snapshot(i, lut)
            setup.cfg.flags |= abi.FLAG_DELTASPH
            run(i, setup, variant, name + " delta-SPH")
    # the artificial stress (176-byte records as well, eigen-solver in the prologue, spacing kernel per target)
    i = golden("stressav_in.snap")
    for variant in (0, 2, 3):
        eng = Engine(abi.setup_from_snapshot(i, lut), len(i["mass"]))
        eng.set_variant(variant)
        eng.upload_state(i, STATE_IN + ("wp",))
        st = eng.integrate()
        print("stressav_in artificial stress variant", variant, "pairs", st.pair_count)
        eng.close()
    setup = workloads.make_setup(len(small["mass"]), solid=True)
    setup.cfg.flags |= abi.FLAG_STRESS_AV
    eng = Engine(setup, len(small["mass"]))
    eng.upload_state(small, names)
    eng.upload_state({"wp": 0.25 / np.pi / small["pos"][:, 3] ** 3}, ["wp"])
    dts, _, st = eng.run_pc(4, 0.01, 10.0)
    print("run_pc 4 steps, artificial stress, pairs", st.pair_count)
    eng.close()
    for solid in (True, False):
        st0 = small if solid else workloads.basalt_sphere_state(6000, 5.0e4, solid=False)
        setup = workloads.make_setup(len(st0["mass"]), solid=solid)
        setup.cfg.flags |= abi.FLAG_DELTASPH
        eng = Engine(setup, len(st0["mass"]))
        eng.upload_state(st0, [k for k in names if k in st0])
        dts, _, st = eng.run_pc(4, 0.01, 10.0)
        print("run_pc 4 steps, delta-SPH, solid", solid, "pairs", st.pair_count)
        eng.close()



def section_gravity_lattice():
    # self-gravity (radix tree, bottom-up moments with arrival counters, warp-wide walk in shared memory), alone and inside
    # batched steps, and the lattice generator
    from opensph_b200.engine import lattice_count, make_lattice  # noqa: E402
    grav_lut = abi.gravity_table_cubic_spline(40000)
    for theta, order, leaf in ((0.5, 3, 0), (0.8, 2, 4), (0.0, 3, 0)):
        setup = workloads.make_setup(len(small["mass"]), solid=True)
        eng = Engine(setup, len(small["mass"]))
        eng.upload_state(small, names)
        eng.gravity_configure(theta, order, abi.GRAVITY_CONSTANT, grav_lut, 2.0, leaf)
        gs = eng.gravity_eval()
        dts, _, st = eng.run_pc(3, 0.01, 10.0)
        print("gravity theta", theta, "order", order, "leaf", leaf, "node interactions", gs.approximated, "ranges", gs.exact, "groups", gs.groups)
        eng.close()
    lat = make_lattice(3000, 5.0e4, (1.0e5, 0.0, -2.0e4))
    m = lattice_count(lat)
    eng = Engine(workloads.make_setup(m, solid=False), m)
    assert eng.lattice_generate(lat) == m
    print("lattice", m, "particles, mass sum", float(eng.download("MASS").sum()))
    eng.close()


# SMOKE_ONLY=deltasph (or pairs / steps / gravity_lattice) runs one section
only = os.environ.get("SMOKE_ONLY")
for sec in (section_pairs, section_steps, section_deltasph, section_gravity_lattice):
    if not only or sec.__name__ == "section_" + only:
        sec()
print("SANITIZER SMOKE DONE")
