"""Initial conditions on the device: the hexagonal packing of a spherical body with its smoothing lengths
and masses, against the plain-C oracle, which is pinned bit for bit on the Storage the reference's InitialConditions built
(tests/golden/preset_in.snap, and live runs when oracle/_ref is present)."""
import os
import tempfile

import numpy as np
import pytest

from conftest import golden, have_ref, run_ref
from opensph_b200 import abi, workloads
import oracle_port as op


def test_oracle_lattice_is_the_reference_lattice():
    g = golden("preset_in.snap")  # InitialConditions::addMonolithicBody(SphericalDomain(0, 5e4), PARTICLE_COUNT = 500)
    pos, mass = op.hexagonal_sphere(500, (0.0, 0.0, 0.0), 5.0e4)
    assert np.array_equal(pos, g["pos"]) and np.array_equal(mass, g["mass"])


@pytest.mark.skipif(not have_ref(), reason="needs the compiled reference (oracle/_ref)")
def test_oracle_lattice_against_live_reference():
    with tempfile.TemporaryDirectory() as tmp:
        i, _ = run_ref(tmp, ["--config", "preset", "--n", 60000, "--no-lut"])
    pos, mass = op.hexagonal_sphere(60000, (0.0, 0.0, 0.0), 5.0e4)
    assert np.array_equal(pos, i["pos"]) and np.array_equal(mass, i["mass"])


def test_bench_workload_lattice_is_the_reference_lattice():
    # bench.py builds its particles with opensph_b200/workloads.py (numpy): the same lattice without the centring shift
    # (coordinates lower + k * step instead of the reference's running sums: equal to rounding)
    pos, mass = op.hexagonal_sphere(20000, (0.0, 0.0, 0.0), 5.0e4, centred=False)
    st = workloads.basalt_sphere_state(20000)
    assert len(st["mass"]) == len(mass)
    assert np.abs(st["pos"] - pos).max() <= 1e-9 * 5.0e4
    assert np.abs(st["mass"] - mass).max() <= 1e-12 * mass.max()


def _device_lattice(n, radius, centre, centred, eta=1.3, rho0=2700.0, flag=3):
    from opensph_b200.engine import Engine, lattice_count, make_lattice
    lat = make_lattice(n, radius, centre, eta, rho0, centred, flag)
    m = lattice_count(lat)
    setup = workloads.make_setup(m, solid=False)
    with Engine(setup, m) as eng:
        assert eng.lattice_generate(lat) == m
        return eng.download_state(["pos", "vel", "acc", "mass", "flag"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,centre", [(500, (0.0, 0.0, 0.0)), (100000, (1.4e5, -3.0e3, 2.5e4))])
def test_gpu_lattice_matches_oracle(n, centre):
    radius = 5.0e4
    # without the centring shift the device's lattice points are the reference's bit for bit
    raw = _device_lattice(n, radius, centre, False)
    pos, mass = op.hexagonal_sphere(n, centre, radius, centred=False)
    assert len(raw["mass"]) == len(mass)
    assert np.array_equal(raw["pos"], pos)
    # the normalisation divides by the sum of all h^3: the reference adds them one after the other (error up to ~N eps),
    # the device as a tree
    mtol = 4.0 * len(mass) * 2.2e-16
    assert np.abs(raw["mass"] - mass).max() <= mtol * mass.max()
    # with it (BodySettings default) the shift is a sum over all particles: equal to rounding
    got = _device_lattice(n, radius, centre, True)
    pos, mass = op.hexagonal_sphere(n, centre, radius, centred=True)
    assert np.abs(got["pos"] - pos).max() <= 1e-12 * radius
    assert np.array_equal(got["pos"][:, 3], pos[:, 3])
    assert np.abs(got["mass"] - mass).max() <= mtol * mass.max()
    assert np.all(got["flag"] == 3) and not got["vel"].any() and not got["acc"].any()


@pytest.mark.gpu
def test_gpu_lattice_at_bench_size():
    """configs[3]: the 10 M-particle sphere generated on the device -- count, total mass and centre of mass."""
    got = _device_lattice(10_000_000, 5.0e4, (0.0, 0.0, 0.0), True)
    n = len(got["mass"])
    assert n == 10625058
    volume = 4.0 / 3.0 * np.pi * 5.0e4 ** 3
    assert abs(got["mass"].sum() - 2700.0 * volume) <= 1e-11 * 2700.0 * volume
    assert np.abs(got["pos"][:, :3].mean(axis=0)).max() <= 1e-6
    assert np.linalg.norm(got["pos"][:, :3], axis=1).max() <= 5.0e4 * (1 + 1e-3)
