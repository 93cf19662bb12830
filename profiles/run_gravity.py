"""Times the device self-gravity (sphgpu_gravity_eval) on the bench's basalt sphere and measures its error against the exact
sums on a sample of targets computed by the device's own exact mode at a smaller size. Run under gpurun:
    python profiles/run_gravity.py 1000000 10000000 > gpurun_out/r02_gravity.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensph_b200 import abi, workloads  # noqa: E402
from opensph_b200.engine import Engine  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1000000]
    grav_lut = abi.gravity_table_cubic_spline(40000)
    out = []
    for n_target in sizes:
        st = workloads.basalt_sphere_state(n_target, solid=False)
        n = len(st["mass"])
        setup = workloads.make_setup(n, solid=False)
        with Engine(setup, n) as eng:
            eng.upload("POSITION", 0, st["pos"])
            eng.upload("MASS", 0, st["mass"])
            rec = {"particles": n}
            for theta, order in ((0.5, 3), (0.8, 3), (0.5, 0)):
                eng.gravity_configure(theta, order, abi.GRAVITY_CONSTANT, grav_lut, 2.0, 20)
                eng.gravity_eval()
                ms = []
                for _ in range(3):
                    s = eng.gravity_eval()
                    ms.append(s.gpu_ms)
                acc = eng.download("POSITION", 2)[:, :3]
                key = f"theta{theta}_order{order}"
                rec[key] = {"ms": float(np.median(ms)), "node_interactions": int(s.approximated), "exact_ranges": int(s.exact), "groups": int(s.groups),
                            "acc_abs_mean": float(np.abs(acc).mean())}
                if n <= 300000:
                    eng.gravity_configure(0.0, 3, abi.GRAVITY_CONSTANT, grav_lut, 2.0, 20)
                    t0 = time.time()
                    eng.gravity_eval()
                    exact = eng.download("POSITION", 2)[:, :3]
                    rec[key]["rms_err_vs_exact"] = float(np.sqrt(((acc - exact) ** 2).sum() / (exact ** 2).sum()))
                    rec["exact_mode_s"] = time.time() - t0
            out.append(rec)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
