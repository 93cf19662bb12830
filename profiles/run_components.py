"""Timing of the connected-component search at the bench size: sphgpu_find_components on the device against
Post::findComponents of the unmodified reference (oracle/_ref/sph_ref components, one host thread -- the reference's flood is
sequential). Run on the GPU box: python profiles/run_components.py > gpurun_out/r02_components.json"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensph_b200 import workloads  # noqa: E402
from opensph_b200.engine import Engine  # noqa: E402

out = {"what": "Post::findComponents, ComponentFlag::OVERLAP", "device": [], "reference": []}
for n_target in (1_000_000, 10_000_000):
    state = workloads.basalt_sphere_state(n_target, 5.0e4, solid=False)
    n = len(state["mass"])
    rng = np.random.default_rng(3)
    pos = state["pos"].copy()
    pos[:, :3] += rng.uniform(-0.25, 0.25, (n, 3)) * pos[:, 3:4]   # the jitter of the golden vectors
    with Engine(workloads.make_setup(n, solid=False), n) as eng:
        eng.upload_state({"pos": pos}, ["pos"])
        for radius in (1.0, 0.62, 0.55):
            eng.find_components(radius)
            dt = 1e30
            for _ in range(3):  # best of three calls (the buffers are allocated and freed inside the call)
                t0 = time.perf_counter()
                idx, count, sweeps = eng.find_components(radius)
                dt = min(dt, time.perf_counter() - t0)
            out["device"].append({"particles": n, "radius": radius, "components": count, "largest": int(np.bincount(idx).max()),
                                  "sweeps": sweeps, "seconds": dt, "note": "wall clock of the C-ABI call: cell list + sweeps + indices to the host"})
ref = os.path.join(ROOT, "oracle", "_ref", "sph_ref")
if os.path.exists(ref):
    for radius in (1.0, 0.62):
        r = subprocess.run([ref, "components", "--config", "collision_preset", "--n", "1000000", "--jitter", "3", "--radius", str(radius),
                            "--out", "/dev/null"], capture_output=True, text=True)
        try:
            out["reference"].append(dict(json.loads(r.stdout.strip().splitlines()[-1]), radius=radius))
        except Exception as e:  # noqa: BLE001
            out["reference"].append({"error": str(e), "stderr": r.stderr[-300:]})
print(json.dumps(out, indent=1))
