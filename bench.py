#!/usr/bin/env python
"""bench.py -- SPH particle-updates/s of the collision-preset step (find + derivatives + integrate + dt criteria).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref/sph_ref)

One step = one PredictorCorrector step of BASELINE.json's collision preset (AsymmetricSolver terms: pressure + solid
stress + standard AV + continuity + adaptive h + correction tensor; Tillotson/von Mises/Grady-Kipp basalt) on a
synthetic hexagonal-lattice basalt sphere. `value` keeps the state resident in HBM; `e2e` moves the state host ->
device and back every step through the same C-ABI calls. With N > 1 ranks the sphere is cut into N slabs of equal
particle count (strong scaling), ghost layers are exchanged over NCCL every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SPH particle-updates/s (find+derivs+integrate, collision preset)"
UNIT = "particle-updates/s"
# Algorithmic HBM bytes per particle-update, SURVEY.md section 8(d), solid + correction tensor:
BYTES_FULL_STEP = 1412.0      # full PredictorCorrector step
BYTES_INTEGRATE = 628.0       # find + derivatives (integrate() only)
FP64_INSTR_PER_PAIR = 97.0     # FP64 instructions of the pair body per neighbour pair (SASS count)
BYTES_PAIR_KERNEL = 392.0     # dominant kernel, itemised in DESIGN.md section 3 (sorted record + epilogue inputs in, derivatives out)
DT_INITIAL, DT_MAX = 0.01, 10.0  # TIMESTEPPING_INITIAL_TIMESTEP / MAX_TIMESTEP of the collision preset (oracle/ref_driver.cpp)


def read_traffic():
    """DRAM bytes per particle of the pair kernel from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md): NVML from a thread every
    20 ms (the timed region of the default run is a fraction of a second), nvidia-smi as the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self.nvml, self.handle, self.running, self.thread, self.max_mhz = None, None, False, None, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, handle

    def _sample_nvml(self):
        p, h = self.nvml, self.handle
        mhz = float(p.nvmlDeviceGetClockInfo(h, p.NVML_CLOCK_SM))
        try:
            bits = int(p.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            bits = int(p.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        self.samples.append((mhz, bits))

    def _loop(self):
        while self.running:
            try:
                self._sample_nvml()
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            try:
                self._sample_nvml()  # the GPU is still busy / boosted when the timed region has just ended
            except Exception:
                pass
            self.running = False
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            sm = sorted(m for m, _ in self.samples)
            reasons = sorted({name for _, bits in self.samples for mask, name in self.REASON_BITS if bits & mask})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


def run_reference(args, n_sample: int):
    """Times the unmodified reference (AsymmetricSolver + KdTree + PredictorCorrector on all host threads) on a bounded
    sample of the same configuration. Returns the JSON fields of the reference's run."""
    exe = os.path.join(ROOT, "oracle", "_ref", "sph_ref")
    if not os.path.exists(exe):
        return None
    cmd = [exe, "bench", "--config", "preset", "--n", str(n_sample), "--steps", str(args.steps), "--warmup",
           str(max(args.warmup, 1)), "--threads", "0"]
    if getattr(args, "dt_mode", "criteria") == "fixed":
        cmd += ["--fixed-dt", "1e-6"]
    out = subprocess.check_output(cmd, text=True)
    return json.loads(out.strip().splitlines()[-1])


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the full configuration whenever the whole run fits in about five minutes of host time (measured: 0.82 us per
    # particle and step on 16 threads), otherwise the largest lattice that does
    budget_particles = int(300.0 / (0.82e-6 * max(args.steps + max(args.warmup, 1), 1)))
    n_sample = args.n if (args.full_reference or args.n <= budget_particles) else max(budget_particles, 100_000)
    t0 = time.time()
    r = run_reference(args, n_sample)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/sph_ref was not built (needs /root/reference)"}))
        return
    value = r["particle_updates_per_s"]
    sample = (f"collision preset, {r['particles']} particles (hex lattice, eta 1.3, {r['neigh_mean']:.1f} mean neighbours), "
              f"{args.steps} PredictorCorrector steps, AsymmetricSolver + KdTree, {r['threads']} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"collision_preset_{args.n}", "sample_particles": r["particles"], "finder": "kd_tree",
                   "threads": r["threads"], "same_config": bool(n_sample == args.n),
                   "dt": "chosen by the preset's criteria (Courant + divergence)" if args.dt_mode == "criteria" else 1e-6},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }))


STATE_NAMES = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "damage", "ddamage", "reduce",
               "eps_min", "m_zero", "growth", "n_flaws", "flag")
STEP_INPUTS = ("pos", "vel", "rho", "u", "S", "damage")     # what a host-resident Storage hands over every step
STEP_OUTPUTS = ("pos", "vel", "rho", "u", "S", "damage")    # the advanced state read back


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", dest="n", type=int, default=10_000_000, help="target particle count of the lattice (configs[3])")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gravity", action="store_true", help="skip the self-gravity figure reported beside the headline")
    ap.add_argument("--fluid", action="store_true", help="fluid-only terms (BASELINE configs[4])")
    ap.add_argument("--terms", choices=["preset", "balsara", "xsph", "deltasph", "stressav"], default="preset",
                    help="optional equation terms on top of the preset (other-workload lines, single GPU; the headline is 'preset')")
    ap.add_argument("--dt-mode", choices=["criteria", "fixed"], default="criteria",
                    help="criteria: the step the preset's criteria choose (particles move); fixed: dt = 1e-6 (static lattice)")
    ap.add_argument("--list-skin", type=float, default=None, help="skin of the candidate-list reuse (0: rebuild every step)")
    ap.add_argument("--full-reference", action="store_true", help="--impl reference at the full particle count (minutes)")
    ap.add_argument("--weak", action="store_true", help="--particles is the count PER GPU (weak scaling, BASELINE configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    from opensph_b200 import workloads
    from opensph_b200.engine import Engine
    from opensph_b200 import decomp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solid = not args.fluid
    if args.dt_mode == "fixed":
        dt0 = dt_max = 1.0e-6  # tiny fixed step: the lattice stays intact, every step does the same work
    else:
        dt0, dt_max = DT_INITIAL, DT_MAX  # first step, then whatever the preset's criteria choose (same as sph_ref bench)
    n_lattice = args.n * world if args.weak else args.n

    # ---- workload -------------------------------------------------------------------------------------------
    dom = decomp.SlabDomain(n_lattice, world, rank, solid=solid)
    state = dom.generate_owned()
    n_owned = len(state["mass"])
    setup = workloads.make_setup(n_owned, solid=solid)
    if args.terms != "preset":
        if world > 1:
            raise SystemExit("--terms needs a single GPU: these terms read results of the previous evaluation, which ghosts do not carry")
        abi_mod = __import__("opensph_b200").abi
        setup.cfg.flags |= {"balsara": abi_mod.FLAG_BALSARA, "xsph": abi_mod.FLAG_XSPH, "deltasph": abi_mod.FLAG_DELTASPH,
                            "stressav": abi_mod.FLAG_STRESS_AV}[args.terms]
    eng = Engine(setup, n_owned, capacity=dom.capacity(n_owned), device=local)
    eng.set_variant(args.variant)
    if args.list_skin is not None:
        eng.set_list_skin(args.list_skin)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.upload_state(state, STATE_NAMES)
    if args.terms == "stressav":  # StressAV::create: W(h, h) of the initial smoothing lengths (cubic spline: 1 / (4 pi h^3))
        eng.upload_state({"wp": 0.25 / np.pi / state["pos"][:, 3] ** 3}, ["wp"])
    halo = decomp.HaloExchange(dom, eng, state) if world > 1 else None
    n_total = dom.total_particles(n_owned)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    cur = {"dt": dt0}  # the step the next call uses: fed back from the criteria

    def one_step():
        dt = cur["dt"]
        if halo is None:
            r = eng.step_pc(dt, dt_max)
        elif halo.native:
            # predict -> NCCL halo exchange -> integrate -> correct -> criteria -> allreduce(min dt), one host sync
            r = eng.step_pc_mgpu(dt, dt_max)
        else:
            # multi-GPU without the native path: ghosts must carry the PREDICTED state, so the step is issued in its parts
            eng.predict(dt)
            halo.exchange()
            st = eng.integrate()
            eng.correct(dt)
            new_dt, crit = eng.compute_timestep(dt_max)
            t = torch.tensor([new_dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)  # global time step = min over ranks
            r = (float(t.item()), crit, st)
        cur["dt"] = r[0]
        return r

    for _ in range(args.warmup):
        one_step()

    # ---- timed region: state resident in HBM -------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pair_ms, launches, timings, halo_ms, pair_parts = 0.0, 0, np.zeros(4), 0.0, np.zeros(3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    batched = halo is None or halo.native  # the K steps are queued back to back, dt stays on the device (sphgpu_run_pc)
    builds0 = eng.list_stats()[0]
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    if batched:
        dts, _, st = eng.run_pc(args.steps, cur["dt"], dt_max)
        cur["dt"] = float(dts[-1])
        launches += st.kernel_launches * args.steps
    else:
        for _ in range(args.steps):
            _, _, st = one_step()
            launches += st.kernel_launches
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    list_builds = eng.list_stats()[0] - builds0  # steps of the timed region that rebuilt cell list + candidate lists
    neigh_mean = st.neigh_mean
    pairs_per_step = float(st.pair_count)
    dt_last = cur["dt"]

    # ---- phase breakdown (not timed above): the same K steps again, one call each, CUDA events of every step -------
    # (with list reuse the steps differ: most skip the cell-list / unit / candidate-list build)
    phase_builds0 = eng.list_stats()[0]
    for _ in range(args.steps):
        one_step()
        tm = eng.last_timings()
        timings += tm
        pair_ms += tm[2]
        pair_parts += eng.last_pair_timings()
        if halo is not None:
            halo_ms += eng.last_halo_ms()
    phase_builds = eng.list_stats()[0] - phase_builds0
    fp64_peak = eng.measure_fp64_peak() if rank == 0 else 0.0

    # ---- parity fields: invariants that must not depend on the number of ranks -----------------------------------
    # global pair count of the last step and a checksum of the accelerations: sum_i w_i a_i over all owned particles, with
    # weights in [0.5, 1.5) keyed by the particle's INITIAL lattice position (the same on every decomposition)
    acc = eng.download_state(["acc"])["acc"][:n_owned, :3]
    h_lat = dom.h / workloads.BASALT["eta"]
    wgt = 0.5 + workloads._hash01(state["pos"][:n_owned, :3], 0.05 * h_lat, 77, 3)
    mass = state["mass"][:n_owned]
    par = torch.tensor([float(st.pair_count)] + [float(np.sum(wgt * acc[:, k])) for k in range(3)] +
                       [float(np.sum(np.abs(acc[:, k]))) for k in range(3)] +
                       [float(np.sum(mass * acc[:, k])) for k in range(3)] + [float(np.sum(mass * np.abs(acc[:, k]))) for k in range(3)],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(par, op=dist.ReduceOp.SUM)
    par = par.tolist()

    tsec = torch.tensor([max(wall, dev_s)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
    step_s = float(tsec.item()) / args.steps
    value = n_total / step_s

    # ---- e2e: host-resident state, H2D + step + D2H every step ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # The call sequence of a driver that keeps its Storage on the host: every step queues the host -> device copies of
        # the state (pinned memory), runs the step, queues the device -> host copies of the advanced state. The copies
        # out of step k leave on a second stream while the copies in of step k + 1 arrive (PCIe is full duplex); all of it
        # is inside the timed region, which ends with the last byte on the host (sphgpu_transfer_sync).
        fields = __import__("opensph_b200").abi.SNAPSHOT_FIELDS
        pinned_in = {k: torch.from_numpy(np.ascontiguousarray(state[k])).pin_memory() for k in STEP_INPUTS if k in state}
        pinned_out = {k: torch.empty_like(v).pin_memory() for k, v in pinned_in.items()}
        np_in = {k: v.numpy() for k, v in pinned_in.items()}
        np_out = {k: v.numpy() for k, v in pinned_out.items()}
        h2d = sum(v.numel() * v.element_size() for v in pinned_in.values())
        d2h = sum(v.numel() * v.element_size() for v in pinned_out.values())

        def e2e_step():
            for k, v in np_in.items():
                eng.upload_async(fields[k][0], fields[k][1], v)
            one_step()
            for k, v in np_out.items():
                eng.download_async(fields[k][0], fields[k][1], v)
            eng.download_batch_end()

        e2e_step()
        eng.transfer_sync()
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(3, min(args.steps, 8))
        for _ in range(k_e2e):
            e2e_step()
        eng.transfer_sync()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total / (float(te.item()) / k_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": k_e2e,
               "what": "sphgpu_upload_async x%d -> step -> sphgpu_download_async x%d per step, pinned host buffers; the download of step k "
                       "overlaps the upload of step k+1" % (len(np_in), len(np_out))}

    # Self-gravity of the same particles: not part of the headline metric (BASELINE's config has gravity
    # off), reported beside it -- one evaluation = keys + sort + tree + moments + walk, at the GUI preset's
    # opening angle 0.8 and at the library default 0.5, octupoles, cubic-spline softening.
    gravity = None
    if world == 1 and not args.no_gravity:
        try:
            grav_lut = __import__("opensph_b200").abi.gravity_table_cubic_spline(40000)
            gravity = {"unit": "ms per evaluation", "particles": int(n_owned)}
            for theta in (0.8, 0.5):
                eng.gravity_configure(theta, 3, __import__("opensph_b200").abi.GRAVITY_CONSTANT, grav_lut, 2.0, 20)
                eng.gravity_eval()
                ms = sorted(eng.gravity_eval().gpu_ms for _ in range(3))[1]
                st = eng.gravity_last_stats()
                gravity[f"opening_angle_{theta}"] = {"ms": float(ms), "node_interactions": int(st.approximated), "exact_ranges": int(st.exact),
                                                     "groups": int(st.groups)}
            eng.gravity_off()
        except Exception as e:
            gravity = {"failed": str(e)}

    # Connected components of the same particles (Post::findComponents): reported beside the headline as well.
    components = None
    if world == 1 and not args.no_gravity:
        try:
            eng.find_components(1.0)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                _, ccount, csweeps = eng.find_components(1.0)
                best = min(best, time.perf_counter() - t0) if best is not None else time.perf_counter() - t0
            components = {"unit": "ms per call", "radius": 1.0, "ms": 1.0e3 * best, "components": int(ccount), "sweeps": int(csweeps),
                          "particles": int(n_owned),
                          "what": "sphgpu_find_components through the C ABI, wall clock: cell list + label-propagation sweeps + indices to the host"}
        except Exception as e:
            components = {"failed": str(e)}

    per_rank = None
    if world > 1:  # per-rank device-time breakdown (ms per step): grid, prologue, pair kernel, rest, of which halo exchange
        mine = torch.tensor(list(timings / args.steps) + [halo_ms / args.steps, float(n_owned)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(x), 4) for x in r.tolist()] for r in allr]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = read_peaks()
    prof = read_traffic() if solid else None
    # the dominant kernel is k_pair_sum (FP64 pair sums); the other variants run everything in one kernel
    sum_ms = pair_parts[2] if pair_parts[2] > 0 else pair_ms
    pair_s = sum_ms * 1e-3 / args.steps
    bytes_pair = (BYTES_PAIR_KERNEL if solid else 222.0) * n_owned
    achieved = bytes_pair / pair_s / 1e9 if pair_s > 0 else 0.0
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": ("fluid_" if args.fluid else "collision_preset_") + str(n_lattice) + ("" if args.terms == "preset" else "+" + args.terms),
                   "particles": int(n_total),
                   "particles_per_gpu": int(n_owned), "mean_neighbours": round(float(neigh_mean), 2), "integrator": "predictor_corrector",
                   "dt": ("chosen by the preset's criteria (Courant + divergence), initial %g, max %g; last step %.4g s" % (dt0, dt_max, dt_last))
                   if args.dt_mode == "criteria" else dt0,
                   "list_reuse": {"skin": args.list_skin if args.list_skin is not None else 0.03, "builds_in_timed_steps": int(list_builds),
                                  "note": "cell list / work units / candidate lists are rebuilt when the device-side displacement "
                                          "check says so; every step applies the exact neighbour predicate"},
                   "l2": "inputs larger than L2 (state %.1f GB per GPU)" % (n_owned * 480 / 1e9),
                   "decomposition": ("z-slabs of equal particle count, halo exchange inside the step: " + getattr(halo, "transport", "nccl point-to-point"))
                   if world > 1 else "single domain",
                   "stepping": "sphgpu_run_pc: K steps queued back to back, time step fed back on the device, one host sync" if batched
                   else "one call and one host sync per step",
                   "pair_variant": args.variant},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "parity": {"pair_count_global": int(round(par[0])), "acc_checksum": [par[1], par[2], par[3]],
                   "acc_abs_sum": [par[4], par[5], par[6]],
                   "momentum_residual": [abs(par[7 + k]) / max(par[10 + k], 1e-300) for k in range(3)],
                   "momentum_note": "|sum_i m_i a_i| / sum_i |m_i a_i| per axis over ALL particles: the pair forces are antisymmetric, so the "
                                    "evaluation conserves linear momentum to rounding at any size (a size-independent check of the pair sums)",
                   "note": "after warm-up + 2 x steps PredictorCorrector steps; must agree across --gpus N (pair count exactly, checksum to "
                           "1e-10 of acc_abs_sum)"},
        "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": "k_pair_sum (list-driven FP64 pair sums + finalizers)" if pair_parts[2] > 0 else
                     "k_pair (fused neighbour search + pair sums + finalizers)",
                     "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " HBM copy GB/s", "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (prof["bytes_per_particle"] * n_owned if prof else None),
                     "traffic_source": (prof["source"] if prof else None),
                     "fp64_pipe_active_pct_ncu": (prof["fp64_pipe_active_pct"] if prof else None),
                     "kernel_ms": pair_s * 1e3, "kernel_share_of_step": pair_s / step_s,
                     "step_hbm_frac": BYTES_FULL_STEP * (n_owned / step_s) / 1e9 / peak,
                     "fp64": {"peak_fma_per_s": fp64_peak, "peak_tflops": 2e-12 * fp64_peak, "kind": "measured live (DFMA loop, all SMs)",
                              "achieved_fp64_instr_per_s": (FP64_INSTR_PER_PAIR * pairs_per_step / pair_s if (solid and pair_s > 0) else None),
                              "frac": (FP64_INSTR_PER_PAIR * pairs_per_step / pair_s / fp64_peak if (solid and pair_s > 0 and fp64_peak > 0) else None)},
                     "note": "pair kernel is FP64-pipe / latency bound, see DESIGN.md; step_hbm_frac uses SURVEY 8(d)'s 1412 B/particle"},
        "integrate_only": {"value": n_owned / (max(timings[0] + timings[1] + timings[2], 1e-9) * 1e-3 / args.steps), "unit": UNIT,
                           "note": "find + derivatives + finalize of this rank (ISolver::integrate alone), device time"},
        "phase_ms": {"note": "mean over %d steps run one call at a time after the timed region; %d of them rebuilt the lists" % (args.steps, phase_builds),
                     "grid_build": timings[0] / args.steps, "prologue_pack": timings[1] / args.steps,
                     "pair_stage": timings[2] / args.steps, "integrator_and_criteria": timings[3] / args.steps,
                     "integrator_and_criteria_in_timed_region": max(step_s * 1e3 - (timings[0] + timings[1] + timings[2]) / args.steps, 0.0) if world == 1 else None,
                     "integrator_note": "single calls run k_predict, k_correct, k_criteria; the timed region (sphgpu_run_pc) runs k_criteria and one fused "
                                        "k_correct_predict per step (corrector of step s + predictor of step s+1: the state is read and written once)",
                     "pair_stage_parts": {"units_and_lane_order": pair_parts[0] / args.steps, "k_pair_lists": pair_parts[1] / args.steps,
                                          "k_pair_sum": pair_parts[2] / args.steps}},
    }
    if gravity is not None:
        out["gravity"] = gravity
    if components is not None:
        out["components"] = components
    if per_rank is not None:
        out["per_rank_ms"] = {"columns": ["grid_build", "prologue_pack", "pair_kernel", "rest", "halo_exchange_in_rest", "owned_particles"],
                              "rows": per_rank}
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference(argparse.Namespace(steps=2, warmup=1, n=args.n, dt_mode=args.dt_mode), min(args.n, 1_000_000))
            if r is not None:
                out["cpu_baseline"] = {
                    "value": r["particle_updates_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                    "sample": f"collision preset, {r['particles']} particles, 2 PredictorCorrector steps after 1 warm-up, "
                              f"AsymmetricSolver + KdTree on {r['threads']} host threads (oracle/_ref/sph_ref)"}
            exe = os.path.join(ROOT, "oracle", "_ref", "sph_ref")
            if gravity is not None and "failed" not in gravity and os.path.exists(exe):
                # the reference's own Barnes-Hut (build + evalSelfGravity) on a 1 M-particle sample, all host threads
                import tempfile
                with tempfile.TemporaryDirectory() as tmp:
                    line = subprocess.check_output([exe, "gravity", "--config", "preset", "--n", str(min(args.n, 1_000_000)), "--gravity", "bh",
                                                    "--theta", "0.5", "--order", "3", "--leaf", "20", "--no-lut", "--out",
                                                    os.path.join(tmp, "g.snap")], text=True)
                rg = json.loads(line.strip().splitlines()[-1])
                gravity["cpu_reference"] = {"particles": rg["particles"], "opening_angle": 0.5, "ms": 1.0e3 * rg["seconds"], "threads": rg["threads"],
                                            "what": "BarnesHut::build + evalSelfGravity of the unmodified reference (oracle/_ref/sph_ref gravity)"}
        except Exception as e:  # the baseline is reported, never required
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
