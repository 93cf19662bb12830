"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_mgpu_parity.py

Every rank owns one x-slab, exchanges ghosts over NCCL and runs two PredictorCorrector steps through the C ABI; rank 0
additionally runs the whole sphere on its own GPU as a single domain and checks that the union of the ranks' results
equals it (neighbour counts exactly, state and derivatives within the parity tolerance)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from compare import rel_err  # noqa: E402
from opensph_b200 import decomp, workloads  # noqa: E402
from opensph_b200.engine import Engine  # noqa: E402

STATE = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
         "eps_min", "m_zero", "growth", "n_flaws", "flag")
CHECK = ("pos", "vel", "rho", "u", "S", "acc", "du", "drho", "dS", "divv")


def steps(eng, halo, n_steps, dt):
    if os.environ.get("MGPU_BATCHED", "1") == "1" and (halo is None or halo.native):
        # all steps queued back to back, the (global) time step fed back on the device: sphgpu_run_pc on both sides
        eng.run_pc(n_steps, dt, dt)
        return
    for _ in range(n_steps):
        if halo is not None and halo.native and os.environ.get("MGPU_NATIVE", "1") == "1":
            eng.step_pc_mgpu(dt, 1.0e30)  # the whole step inside the library, NCCL on the engine's stream
            continue
        eng.predict(dt)
        if halo is not None:
            halo.exchange()
        eng.integrate()
        eng.correct(dt)


def main():
    n_target = int(os.environ.get("MGPU_PARTICLES", "200000"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = 1.0e-3
    dom = decomp.SlabDomain(n_target, world, rank, radius=5.0e4)
    state = dom.generate_owned()
    n = len(state["mass"])
    eng = Engine(workloads.make_setup(n), n, capacity=dom.capacity(n), device=local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.upload_state(state, STATE)
    halo = decomp.HaloExchange(dom, eng, state)
    steps(eng, halo, 2, dt)
    got = eng.download_state(CHECK + ("ncnt",))
    # gather everything on rank 0
    gathered = [None] * world
    dist.gather_object(got, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        full = workloads.basalt_sphere_state(n_target, 5.0e4)
        # flaw constants are keyed by position (workloads._hash01): both runs see identical inputs
        allp = {k: np.concatenate([g[k] for g in gathered]) for k in CHECK + ("ncnt",)}
        nf = len(full["mass"])
        assert len(allp["ncnt"]) == nf, (len(allp["ncnt"]), nf)
        single = Engine(workloads.make_setup(nf), nf, device=local)
        single.upload_state(full, STATE)
        steps(single, None, 2, dt)
        ref = single.download_state(CHECK + ("ncnt",))

        def key(pos0):
            return np.lexsort((np.round(pos0[:, 0], 0), np.round(pos0[:, 1], 0), np.round(pos0[:, 2], 0)))

        # initial positions identify particles (state has moved by ~dt*v << lattice spacing)
        a, b = key(allp["pos"]), key(ref["pos"])
        if not np.array_equal(allp["ncnt"][a], ref["ncnt"][b]):
            print("FAIL ncnt differs:", int((allp["ncnt"][a] != ref["ncnt"][b]).sum()))
            ok = False
        for k in CHECK:
            e = rel_err(allp[k][a], ref[k][b], 1e-4)
            print(f"{k:6s} rel err {e:.3e}")
            ok &= e <= 1e-9
        print("MGPU PARITY", "OK" if ok else "FAILED", f"world={world} particles={nf} ghosts={halo.g_left + halo.g_right}")
    # ---- re-cut + migration: displace the particles across the cut planes, repartition, compare again ----
    def warp(pos0):
        out = pos0.copy()
        out[:, 2] += 0.2 * 5.0e4 * np.sin(2.0 * np.pi * pos0[:, 0] / 5.0e4)
        return out

    n_before = eng.n
    eng.upload("POSITION", 0, warp(eng.download_state(["pos"])["pos"]))
    _, halo2 = decomp.repartition(dom, eng)
    halo2.exchange()
    eng.integrate()
    got2 = eng.download_state(("pos", "acc", "du", "drho", "dS", "divv", "ncnt"))
    gathered2 = [None] * world
    dist.gather_object(got2, gathered2 if rank == 0 else None, dst=0)
    if rank == 0:
        moved = warp(single.download_state(["pos"])["pos"])
        single.upload("POSITION", 0, moved)
        single.integrate()
        ref2 = single.download_state(("pos", "acc", "du", "drho", "dS", "divv", "ncnt"))
        all2 = {k: np.concatenate([g[k] for g in gathered2]) for k in ref2}
        ok2 = len(all2["ncnt"]) == nf
        if ok2:
            a, b = key(all2["pos"]), key(ref2["pos"])
            ok2 = np.array_equal(all2["ncnt"][a], ref2["ncnt"][b])
            for k in ("acc", "du", "drho", "dS", "divv"):
                e = rel_err(all2[k][a], ref2[k][b], 1e-4)
                print(f"{k:6s} rel err after repartition {e:.3e}")
                ok2 &= e <= 1e-9
        print("MGPU REPARTITION", "OK" if ok2 else "FAILED", f"owned before={n_before} (rank 0) after={[len(g['ncnt']) for g in gathered2]}")
        ok &= bool(ok2)
    # ---- halo guard: grow every smoothing length by 30 %; the fixed send bands (head-room 25 %) are then too narrow. The
    # first step still passes (the guard's h_max bound is the one of the last list build), the next exchange must fail
    # hard on every rank that has a neighbour -- never silently drop cross-rank neighbours.
    from opensph_b200.engine import SphGpuError
    from opensph_b200 import abi
    margin_before = eng.halo_margin()
    pos = eng.download_state(["pos"])["pos"]
    pos[:, 3] *= 1.3
    eng.upload("POSITION", 0, pos)
    raised = False
    try:
        for _ in range(3):
            eng.step_pc_mgpu(1.0e-6, 1.0e-6)
    except SphGpuError as e:
        raised = e.code == abi.E_STATE
    guard_ok = torch.tensor([1 if (raised and margin_before > 0.0) else 0], device="cuda")
    dist.all_reduce(guard_ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU HALO GUARD", "OK" if int(guard_ok.item()) else "FAILED", f"head-room before growing h: {margin_before:.3f}")
        ok &= bool(int(guard_ok.item()))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
