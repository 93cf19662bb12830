"""Connected components (Post::findComponents, core/post/Analysis.cpp:36-75,115-128): integer / index work, so
the bar is bit-exact. Golden vectors come from the reference (sph_ref components: jittered positions AND smoothing lengths, so
the directed relation |r_j - r_i| < h_i * radius is not symmetric); CPU tests pin the oracle's flood restatement on them, the
-m gpu tests the device's label propagation (minimum-index ancestor + pointer jumping, components.cu) against the golden
vectors, against the oracle on clouds with strongly varying h, and at the bench size through size-independent properties."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden
from opensph_b200 import abi, workloads
import oracle_port

CASES = [("components_r07.snap", False), ("components_r09_flag.snap", True), ("components_r10.snap", False),
         ("components_hello4k_r08.snap", False)]


def oracle_components(pos, radius, flag=None):
    lib = oracle_port.lib()
    lib.orc_find_components.restype = C.c_uint32
    pos = np.ascontiguousarray(pos, np.float64)
    idx = np.zeros(len(pos), np.uint32)
    fl = None if flag is None else np.ascontiguousarray(flag, np.uint32)
    count = lib.orc_find_components(pos.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint32(len(pos)), C.c_double(radius),
                                    None if fl is None else fl.ctypes.data_as(C.POINTER(C.c_uint32)), idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return idx, int(count)


def random_cloud(n, seed, h_spread):
    """Points in a unit box with h varying by a factor of up to h_spread (a strongly directed relation)."""
    rng = np.random.default_rng(seed)
    pos = np.zeros((n, 4))
    pos[:, :3] = rng.uniform(0, 1, (n, 3))
    pos[:, 3] = n ** (-1.0 / 3.0) * rng.uniform(1.0, h_spread, n)
    return pos


@pytest.mark.parametrize("name,by_flag", CASES)
def test_oracle_components_match_golden(name, by_flag):
    g = golden(name)
    radius, flags, count = g["comp_params"][:3]
    assert bool(int(flags) & 1) == by_flag
    idx, n = oracle_components(g["pos"], radius, g["flag"] if by_flag else None)
    assert n == int(count)
    assert np.array_equal(idx, g["comp_idx"])


def test_directed_relation_is_exercised_by_the_golden_vectors():
    """With the jittered h some pairs are joined in one direction only, and the flood order matters for them."""
    g = golden("components_r07.snap")
    pos, radius = g["pos"], g["comp_params"][0]
    d2 = ((pos[:, None, :3] - pos[None, :, :3]) ** 2).sum(-1)
    reach = (pos[:, 3] * radius) ** 2
    fwd = d2 < reach[:, None]
    assert (fwd & ~fwd.T).sum() > 10


@pytest.mark.parametrize("seed,spread,radius", [(11, 1.0, 0.9), (12, 3.0, 0.55), (13, 6.0, 0.35)])
def test_min_ancestor_labels_equal_the_flood(seed, spread, radius):
    """The statement the device algorithm rests on (components.cu), checked without a GPU: iterating
    label(p) = min(p, min over q -> p of label(q)) to its fixed point over the DIRECTED relation, and ranking the roots in index
    order, reproduces the sequential flood of the reference -- numbering included."""
    n = 1500
    pos = random_cloud(n, seed, spread)
    d2 = ((pos[:, None, :3] - pos[None, :, :3]) ** 2).sum(-1)
    edge = d2 < ((pos[:, 3] * radius) ** 2)[:, None]   # edge[q, p]: p lies within the reach of q
    assert (edge & ~edge.T).any() or spread == 1.0
    label = np.arange(n)
    for _ in range(n):
        pushed = np.where(edge, label[:, None], n).min(axis=0)   # lowest label among the particles that reach p
        new = np.minimum(label, pushed)
        new = np.minimum(new, new[new])                          # pointer jumping: an ancestor's ancestor is an ancestor
        if np.array_equal(new, label):
            break
        label = new
    roots = np.flatnonzero(label == np.arange(n))
    rank = np.zeros(n, np.int64)
    rank[roots] = np.arange(len(roots))
    ref, count = oracle_components(pos, radius)
    assert count == len(roots) and 1 < count < n
    assert np.array_equal(rank[label], ref)


def engine_with_positions(pos, flag=None):
    from opensph_b200.engine import Engine
    n = len(pos)
    eng = Engine(workloads.make_setup(n, solid=False), n)
    eng.upload_state({"pos": np.ascontiguousarray(pos)}, ["pos"])
    if flag is not None:
        eng.upload_state({"flag": np.ascontiguousarray(flag, np.uint32)}, ["flag"])
    return eng


@pytest.mark.gpu
@pytest.mark.parametrize("name,by_flag", CASES)
def test_gpu_components_match_golden(name, by_flag):
    g = golden(name)
    radius, _, count = g["comp_params"][:3]
    with engine_with_positions(g["pos"], g["flag"]) as eng:
        idx, n, sweeps = eng.find_components(radius, by_flag)
    assert n == int(count)
    assert np.array_equal(idx, g["comp_idx"])
    assert 1 <= sweeps <= 64


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,spread,radius", [(3000, 1, 1.0, 0.9), (3000, 2, 3.0, 0.55), (20000, 3, 2.0, 0.7), (20000, 4, 6.0, 0.35),
                                                  (200000, 5, 2.5, 0.6)])
def test_gpu_components_match_oracle_on_directed_clouds(n, seed, spread, radius):
    pos = random_cloud(n, seed, spread)
    flag = (np.arange(n) % 3).astype(np.uint32)
    for by_flag in (False, True):
        ref, count = oracle_components(pos, radius, flag if by_flag else None)
        with engine_with_positions(pos, flag) as eng:
            idx, m, _ = eng.find_components(radius, by_flag)
        assert m == count and 1 < count < n
        assert np.array_equal(idx, ref)


@pytest.mark.gpu
def test_gpu_components_edge_cases():
    from opensph_b200.engine import SphGpuError
    # one particle; all isolated; all joined; a chain whose only links point "backwards" in index
    one = np.array([[0.5, 0.5, 0.5, 0.1]])
    with engine_with_positions(one) as eng:
        idx, m, _ = eng.find_components(1.0)
        assert m == 1 and idx.tolist() == [0]
        with pytest.raises(SphGpuError):
            eng.find_components(0.0)
    pos = random_cloud(500, 9, 1.0)
    with engine_with_positions(pos) as eng:
        idx, m, _ = eng.find_components(1e-3)
        assert m == 500 and np.array_equal(idx, np.arange(500))
        idx, m, _ = eng.find_components(50.0)
        assert m == 1 and not idx.any()
    n = 400
    chain = np.zeros((n, 4))
    chain[:, 0] = np.arange(n)[::-1]          # particle 0 sits at the far end
    chain[:, 3] = 1.0
    chain[:, 1] = 0.5
    chain[:, 2] = 0.5
    ref, count = oracle_components(chain, 1.5)
    with engine_with_positions(chain) as eng:
        idx, m, sweeps = eng.find_components(1.5)
    assert m == count == 1 and np.array_equal(idx, ref)
    assert sweeps < 40  # pointer jumping: far fewer sweeps than the 400 hops of the chain


@pytest.mark.gpu
def test_gpu_components_at_bench_size():
    """10.6 M particles: the solid sphere is one component at the kernel's reach, every particle is its own below the lattice
    spacing, and two separated spheres give two components numbered by their first particle."""
    from opensph_b200.engine import Engine
    state = workloads.basalt_sphere_state(10_000_000, 5.0e4, solid=False)
    n = len(state["mass"])
    with Engine(workloads.make_setup(n, solid=False), n) as eng:
        eng.upload_state(state, ["pos"])
        idx, m, sweeps = eng.find_components(1.0)
        assert m == 1 and not idx.any()
        idx, m, _ = eng.find_components(0.3)
        assert m == n and np.array_equal(idx, np.arange(n, dtype=np.uint32))
        pos = state["pos"].copy()
        half = pos[:, 0] > 0
        pos[half, 0] += 2.0e4   # pull the sphere apart along x
        eng.upload_state({"pos": pos}, ["pos"])
        idx, m, _ = eng.find_components(1.0)
        assert m == 2
        first = int(idx[0])
        assert first == 0 and np.array_equal(idx == idx[0], half == half[0])
