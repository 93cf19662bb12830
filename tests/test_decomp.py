"""CPU tests of the multi-GPU host logic (world_size 2, gloo): slab decomposition + halo exchange reproduce the
single-domain result. The per-rank compute engine is stood in for by the plain-C oracle so the test needs no GPU; on
the GPU box the same HaloExchange drives libsphgpu through EngineAdapter (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from opensph_b200 import decomp, workloads


def test_cut_planes_split_sphere_evenly():
    cuts = decomp.sphere_cut_planes(1.0, 4)
    assert cuts[0] == -1.0 and cuts[-1] == 1.0 and np.all(np.diff(cuts) > 0)
    pos, _ = workloads.hexagonal_sphere(40000, 1.0)
    counts = np.histogram(pos[:, 0], bins=cuts)[0]
    assert counts.sum() == len(pos)
    assert counts.max() / counts.min() < 1.08  # lattice discreteness at this small N


def test_band_partition_is_stable_and_contiguous():
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 10, 1000)
    perm, nl, nr = decomp.band_partition(x, 0.0, 10.0, 1.5)
    xs = x[perm]
    assert np.all(xs[:nl] < 1.5) and np.all(xs[len(x) - nr:] >= 8.5)
    assert np.all((xs[nl:len(x) - nr] >= 1.5) & (xs[nl:len(x) - nr] < 8.5))
    for seg in (perm[:nl], perm[nl:len(x) - nr], perm[len(x) - nr:]):
        assert np.all(np.diff(seg) > 0)  # stable: original order kept inside each band
    with pytest.raises(ValueError):
        decomp.band_partition(x, 0.0, 10.0, 6.0)


class NumpyEngine:
    """Stand-in for opensph_b200.engine.Engine: slot arrays on the host."""

    def __init__(self, state, capacity):
        self.capacity = capacity
        self.n = len(state["mass"])
        self.n_active = self.n
        self.a = {}
        for k, v in state.items():
            if isinstance(v, np.ndarray) and v.shape[:1] == (self.n,):
                buf = np.zeros((capacity,) + v.shape[1:], v.dtype)
                buf[: self.n] = v
                self.a[k] = buf

    def set_active(self, n):
        self.n_active = n

    def set_particle_count(self, n):
        assert n <= self.capacity
        self.n = self.n_active = n

    def download_state(self, names):
        return {k: self.a[k][: self.n].copy() for k in names if k in self.a}

    def upload_state(self, arrays, names=None, first=0):
        for k in (names if names is not None else arrays.keys()):
            if k in arrays and k in self.a:
                self.a[k][first:first + len(arrays[k])] = arrays[k]

    def upload(self, quantity, order, arr, first=0):
        name = {"FLAG": "flag", "MATERIAL_ID": "matid"}[quantity]
        if name in self.a:
            self.a[name][first:first + len(arr)] = arr

    def download(self, quantity, order=0, first=0, count=None):
        name = {"FLAG": "flag", "MATERIAL_ID": "matid"}[quantity]
        count = self.n - first if count is None else count
        return self.a[name][first:first + count].copy() if name in self.a else np.zeros(count, np.uint32)


class NumpyAdapter:
    def __init__(self, eng):
        self.eng = eng

    def new_buffer(self, doubles):
        import torch
        return torch.empty(max(doubles, 1), dtype=torch.float64)

    def pack(self, fields, first, count, buf):
        import torch
        off = 0
        for name, ncomp in fields:
            buf[off:off + ncomp * count] = torch.from_numpy(
                np.ascontiguousarray(self.eng.a[name][first:first + count]).reshape(-1))
            off += ncomp * count

    def unpack(self, fields, first, count, buf):
        off = 0
        for name, ncomp in fields:
            block = buf[off:off + ncomp * count].numpy()
            self.eng.a[name][first:first + count] = block.reshape(self.eng.a[name][first:first + count].shape)
            off += ncomp * count


def _worker(rank, world, n_target, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle_port import OraclePort
    dist.init_process_group("gloo", init_method=f"file://{tmpdir}/rendezvous", rank=rank, world_size=world)
    dom = decomp.SlabDomain(n_target, world, rank, radius=1.0e3, solid=True)
    state = dom.generate_owned()
    n_owned = len(state["mass"])
    # two "bodies" and two "materials" split by a plane that crosses every slab: ghosts must arrive with their owner's values
    state["flag"] = (state["pos"][:, 0] > 0.0).astype(np.uint32)
    state["matid"] = (state["pos"][:, 1] > 0.0).astype(np.uint32)
    eng = NumpyEngine(state, dom.capacity(n_owned))
    halo = decomp.HaloExchange(dom, eng, state, adapter=NumpyAdapter(eng))
    halo.exchange()
    n_act = halo.n_active
    assert np.array_equal(eng.a["flag"][n_owned:n_act], (eng.a["pos"][n_owned:n_act, 0] > 0.0).astype(np.uint32))
    assert np.array_equal(eng.a["matid"][n_owned:n_act], (eng.a["pos"][n_owned:n_act, 1] > 0.0).astype(np.uint32))
    eng.a["flag"][:n_act] = 0  # (the oracle check below runs the single-body setup)
    snap = {k: v[:n_act].copy() for k, v in eng.a.items()}
    # ghosts: static per-particle constants that only matter for targets get neutral values
    for k, fill in (("reduce", 1.0), ("eps_min", 1.0), ("m_zero", 1.0), ("growth", 0.0)):
        if k in snap:
            snap[k][n_owned:] = fill
    if "n_flaws" in snap:
        snap["n_flaws"][n_owned:] = 1
    setup = workloads.make_setup(n_act, solid=True)
    orc = OraclePort(snap, setup)
    orc.integrate()
    out = {k: orc.a[k][:n_owned] for k in ("pos", "acc", "du", "drho", "dS", "divv", "ncnt")}
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), ghosts=np.array([halo.g_left, halo.g_right]), **out)
    dist.destroy_process_group()


def test_two_rank_halo_exchange_matches_single_domain(tmp_path):
    import torch.multiprocessing as mp
    from oracle_port import OraclePort
    from compare import assert_close

    n_target = 6000
    mp.spawn(_worker, args=(2, n_target, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert parts[0]["ghosts"][1] > 0 and parts[1]["ghosts"][0] > 0
    # single-domain reference on the same lattice (same generator, no x_range)
    full = workloads.basalt_sphere_state(n_target, 1.0e3, solid=True)
    n = len(full["mass"])
    assert sum(len(p["ncnt"]) for p in parts) == n
    # flaw constants are drawn per rank from different RNG streams; they do not enter dv/du/drho/dS
    orc = OraclePort(full, workloads.make_setup(n, solid=True))
    orc.integrate()

    def key(pos):
        return np.lexsort((np.round(pos[:, 0], 6), np.round(pos[:, 1], 6), np.round(pos[:, 2], 6)))

    ref_order = key(orc.a["pos"])
    got = {k: np.concatenate([p[k] for p in parts]) for k in ("pos", "acc", "du", "drho", "dS", "divv", "ncnt")}
    got_order = key(got["pos"])
    assert np.allclose(got["pos"][got_order, :3], orc.a["pos"][ref_order, :3], rtol=0, atol=1e-9)
    assert np.array_equal(got["ncnt"][got_order], orc.a["ncnt"][ref_order])
    for k in ("acc", "du", "drho", "dS", "divv"):
        assert_close(k, got[k][got_order], orc.a[k][ref_order], 1e-10, 1e-4)


def _warp(pos, radius):
    """Smooth displacement that carries particles across the z = 0 cut plane in both directions."""
    out = pos.copy()
    out[:, 2] += 0.25 * radius * np.sin(2.0 * np.pi * pos[:, 0] / radius)
    return out


def _worker_repartition(rank, world, n_target, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle_port import OraclePort
    radius = 1.0e3
    dist.init_process_group("gloo", init_method=f"file://{tmpdir}/rendezvous", rank=rank, world_size=world)
    dom = decomp.SlabDomain(n_target, world, rank, radius=radius, solid=True)
    state = dom.generate_owned()
    state["pos"] = _warp(state["pos"], radius)
    n_before = len(state["mass"])
    # room for everything a rank may receive: this test moves a quarter of the particles
    eng = NumpyEngine(state, 2 * dom.capacity(n_before))
    names = [k for k in decomp.MIGRATE_FIELDS if k in eng.a]
    state, halo = decomp.repartition(dom, eng, names, adapter=NumpyAdapter(eng))
    n_owned = len(state["mass"])
    lo = dom.lo_plane if dom.lo_plane is not None else -np.inf
    hi = dom.hi_plane if dom.hi_plane is not None else np.inf
    assert np.all((state["pos"][:, 2] >= lo) & (state["pos"][:, 2] < hi)), "a particle is outside its slab after migration"
    halo.exchange()
    n_act = halo.n_active
    snap = {k: v[:n_act].copy() for k, v in eng.a.items()}
    for k, fill in (("reduce", 1.0), ("eps_min", 1.0), ("m_zero", 1.0), ("growth", 0.0)):
        if k in snap:
            snap[k][n_owned:] = fill
    if "n_flaws" in snap:
        snap["n_flaws"][n_owned:] = 1
    orc = OraclePort(snap, workloads.make_setup(n_act, solid=True))
    orc.integrate()
    out = {k: orc.a[k][:n_owned] for k in ("pos", "acc", "du", "drho", "dS", "divv", "ncnt")}
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), counts=np.array([n_before, n_owned]), **out)
    dist.destroy_process_group()


def test_repartition_migrates_particles_and_keeps_the_result(tmp_path):
    """re-cut the slabs and migrate. After a displacement that carries many particles across the cut plane,
    repartition() must leave every particle on exactly one rank, inside its slab, the counts balanced, and the
    decomposed evaluation equal to the single-domain one."""
    import torch.multiprocessing as mp
    from oracle_port import OraclePort
    from compare import assert_close

    n_target, radius = 6000, 1.0e3
    mp.spawn(_worker_repartition, args=(2, n_target, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    full = workloads.basalt_sphere_state(n_target, radius, solid=True)
    full["pos"] = _warp(full["pos"], radius)
    n = len(full["mass"])
    owned = [int(p["counts"][1]) for p in parts]
    assert sum(owned) == n
    assert abs(owned[0] - owned[1]) <= 0.02 * n  # re-cut to (nearly) equal counts
    orc = OraclePort(full, workloads.make_setup(n, solid=True))
    orc.integrate()

    def key(pos):
        return np.lexsort((np.round(pos[:, 0], 6), np.round(pos[:, 1], 6), np.round(pos[:, 2], 6)))

    ref_order = key(orc.a["pos"])
    got = {k: np.concatenate([p[k] for p in parts]) for k in ("pos", "acc", "du", "drho", "dS", "divv", "ncnt")}
    got_order = key(got["pos"])
    assert np.allclose(got["pos"][got_order, :3], orc.a["pos"][ref_order, :3], rtol=0, atol=1e-9)
    assert np.array_equal(got["ncnt"][got_order], orc.a["ncnt"][ref_order])
    for k in ("acc", "du", "drho", "dS", "divv"):
        assert_close(k, got[k][got_order], orc.a[k][ref_order], 1e-10, 1e-4)


def test_balanced_cut_planes_single_process():
    rng = np.random.default_rng(3)
    z = np.concatenate([rng.normal(0, 1, 30000), rng.uniform(4, 5, 10000)])  # strongly non-uniform
    cuts = decomp.balanced_cut_planes(z, 4)
    counts = np.histogram(z, bins=cuts)[0]
    assert counts.sum() == len(z)
    assert counts.max() - counts.min() <= 0.01 * len(z)
