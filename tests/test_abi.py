"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/sphgpu.h declares, the ctypes
mirror matches the C struct sizes, and without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from opensph_b200 import abi, engine

HEADER = os.path.join(ROOT, "include", "sphgpu.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"SPHGPU_API\s+[\w\s\*]+?\b(sphgpu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sphgpu.h but not exported by libsphgpu.so"


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sphgpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(sphgpu_config),'
                   ' sizeof(sphgpu_material), sizeof(sphgpu_stats), sizeof(sphgpu_timestep), sizeof(sphgpu_gravity),'
                   ' sizeof(sphgpu_gravity_stats), sizeof(sphgpu_lattice), sizeof(sphgpu_frozen));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(abi.Config), C.sizeof(abi.Material), C.sizeof(abi.Stats), C.sizeof(abi.TimeStep),
                     C.sizeof(abi.Gravity), C.sizeof(abi.GravityStats), C.sizeof(abi.Lattice), C.sizeof(abi.Frozen)]


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(lut):
    snap = golden("hello_in.snap")
    setup = abi.setup_from_snapshot(snap, lut)
    with pytest.raises(engine.SphGpuError) as e:
        engine.Engine(setup, len(snap["mass"]))
    assert e.value.code == abi.E_NO_DEVICE


def test_invalid_setups_are_rejected(lut):
    snap = golden("hello_in.snap")
    n = len(snap["mass"])
    setup = abi.setup_from_snapshot(snap, lut)
    setup.cfg.discretization = 1  # BENZ_ASPHAUG has no GPU implementation -> InvalidSetup, not a silent CPU path
    with pytest.raises(engine.SphGpuError) as e:
        engine.Engine(setup, n)
    assert e.value.code == abi.E_INVALID
    setup = abi.setup_from_snapshot(snap, lut)
    setup.materials[0].end = n - 1  # ranges must cover all particles
    with pytest.raises(engine.SphGpuError) as e:
        engine.Engine(setup, n)
    assert e.value.code == abi.E_INVALID


def test_quantity_ids_match_header():
    text = open(HEADER).read()
    ids = {m.group(1): int(m.group(2)) for m in re.finditer(r"SPHGPU_Q_(\w+)\s*=\s*(\d+)", text)}
    for name, (qid, _, _) in abi.QUANTITIES.items():
        assert ids[name] == qid, name
    assert ids["COUNT"] == len(abi.QUANTITIES)


def test_workload_constants_are_decomposition_independent():
    """Per-particle flaw constants are keyed by position, so a rank generating only its slab draws the same values."""
    from opensph_b200 import workloads
    a = workloads.basalt_sphere_state(4000)
    b = workloads.basalt_sphere_state(4000, x_range=(-1e9, 0.0), total_hint=len(a["mass"]), axis=2)
    m = a["pos"][:, 2] < 0
    assert m.sum() == len(b["mass"])
    assert np.array_equal(a["eps_min"][m], b["eps_min"]) and np.array_equal(a["n_flaws"][m], b["n_flaws"])
