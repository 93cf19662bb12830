"""FrozenParticles boundary condition (core/sph/boundary/Boundary.cpp:203-258): bodies frozen by flag and
particles near / outside a spherical domain. Golden vector: the reference's solver constructed with the boundary condition
(tests/golden/make_golden.sh)."""
import numpy as np
import pytest

from conftest import golden
from compare import assert_close
from opensph_b200 import abi
from oracle_port import OraclePort

FLOOR = 1e-4
STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag")
OUT = ("pos", "acc", "du", "drho", "dS", "ddamage", "divv", "vel")
DOMAIN = ((0.0, 0.0, 0.0), 9.0e4, 0.3)


def _expect_some_of_each(i, o):
    frozen = np.abs(o["acc"][:, :3]).sum(axis=1) == 0
    moved = (o["pos"][:, :3] != i["pos"][:, :3]).any(axis=1)
    assert 20 < frozen.sum() < len(frozen) - 20 and moved.sum() > 5
    assert np.all(frozen[i["flag"] == 1])


def test_oracle_frozen_matches_golden(lut):
    i, o = golden("collision_in.snap"), golden("frozen_out.snap")
    _expect_some_of_each(i, o)
    orc = OraclePort(i, abi.setup_from_snapshot(i, lut))
    orc.integrate()
    orc.frozen(flags=(1,), domain=DOMAIN)
    assert np.array_equal(orc.a["pos"], o["pos"])
    for k in OUT:
        assert_close(k, orc.a[k], o[k], 1e-10, FLOOR)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_gpu_frozen_matches_golden(variant, lut):
    from opensph_b200.engine import Engine
    i, o = golden("collision_in.snap"), golden("frozen_out.snap")
    setup = abi.setup_from_snapshot(i, lut)
    with Engine(setup, len(i["mass"])) as eng:
        eng.set_variant(variant)
        eng.upload_state(i, STATE_IN)
        eng.set_frozen(flags=(1,), domain=DOMAIN)
        eng.integrate()
        got = eng.download_state(list(OUT) + ["ncnt"])
        # batched steps keep applying it: the frozen body does not accelerate
        eng.run_pc(3, 1e-3, 1e-2)
        after = eng.download_state(["vel"])
        eng.set_frozen()
        eng.upload_state(i, STATE_IN)
        eng.integrate()
        free = eng.download_state(["acc"])
    assert np.array_equal(got["ncnt"], o["ncnt"])
    assert np.abs(got["pos"] - o["pos"]).max() <= 1e-15 * 1.0e5
    for k in OUT:
        assert_close(k, got[k], o[k], 1e-10, FLOOR)
    impactor = i["flag"] == 1
    assert np.array_equal(after["vel"][impactor, :3], got["vel"][impactor, :3])
    assert_close("acc without the boundary condition", free["acc"], golden("collision_out.snap")["acc"], 1e-10, FLOOR)
