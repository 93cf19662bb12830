"""Achieved GPU-vs-reference error of one integrate() under the survey's scale floor (1e-6 max|q|) and under the floor the
tests use (1e-4 max|q|), next to the difference between the reference's own two builds (strict IEEE vs upstream fast-math
flags) on the same input. Run under gpurun:  python profiles/error_floors.py > gpurun_out/r02_error_floors.json"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from compare import rel_err  # noqa: E402
from conftest import REF_FAST, REF_STRICT, run_ref  # noqa: E402
from opensph_b200 import abi  # noqa: E402
from opensph_b200.engine import Engine  # noqa: E402

STATE_IN = ("pos", "vel", "mass", "rho", "u", "p", "cs", "S", "damage", "reduce", "eps_min", "m_zero", "growth", "n_flaws", "flag")
KEYS = ("acc", "du", "drho", "dS", "divv", "gradv", "corr")
out = {}
for label, args in (("preset_lattice_200k", ["--config", "preset", "--n", 200000]),
                    ("collision_preset_jitter_100k", ["--config", "collision_preset", "--n", 100000, "--jitter", 3])):
    with tempfile.TemporaryDirectory() as tmp:
        i, o = run_ref(tmp, args, REF_STRICT)
    with tempfile.TemporaryDirectory() as tmp:
        _, f = run_ref(tmp, args, REF_FAST)
    setup = abi.setup_from_snapshot(i)
    with Engine(setup, len(i["mass"])) as eng:
        eng.upload_state(i, STATE_IN)
        eng.integrate()
        got = eng.download_state(list(KEYS) + ["ncnt"])
    rec = {"particles": len(i["mass"]), "neighbour_counts_equal": bool((got["ncnt"] == o["ncnt"]).all())}
    for k in KEYS:
        rec[k] = {"gpu_vs_reference_floor_1e-6": rel_err(got[k], o[k], 1e-6), "gpu_vs_reference_floor_1e-4": rel_err(got[k], o[k], 1e-4),
                  "reference_fast_vs_strict_floor_1e-6": rel_err(f[k], o[k], 1e-6), "reference_fast_vs_strict_floor_1e-4": rel_err(f[k], o[k], 1e-4)}
    out[label] = rec
print(json.dumps(out, indent=1))
