"""One connected-component search of the 10.6 M-particle jittered cloud (the workload of run_components.py), for ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensph_b200 import workloads  # noqa: E402
from opensph_b200.engine import Engine  # noqa: E402

state = workloads.basalt_sphere_state(10_000_000, 5.0e4, solid=False)
n = len(state["mass"])
pos = state["pos"].copy()
pos[:, :3] += np.random.default_rng(3).uniform(-0.25, 0.25, (n, 3)) * pos[:, 3:4]
with Engine(workloads.make_setup(n, solid=False), n) as eng:
    eng.upload_state({"pos": pos}, ["pos"])
    print(eng.find_components(1.0)[1:])
