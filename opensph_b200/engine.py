"""Python binding of the C ABI (include/sphgpu.h) -- thin ctypes layer used by the tests, bench.py and the multi-GPU
driver. It adds no compute of its own and has no fallback: if libsphgpu.so is missing or no CUDA device is usable
it raises.

The call sequence mirrors the reference's (core/run/IRun.cpp:304-329, core/timestepping/TimeStepping.cpp:324-346):
create (AsymmetricSolver ctor) -> upload (Storage) -> integrate / step_pc (ISolver::integrate, ITimeStepping::step)
-> download.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from . import abi

# (SPHGPU_LIB selects another build of the same library, e.g. one compiled with different tile parameters, for A/B runs)
_LIB_PATH = os.environ.get("SPHGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsphgpu.so")
_lib: Optional[C.CDLL] = None


class SphGpuError(RuntimeError):
    """Non-zero status from libsphgpu (the C++ wrapper maps the same codes to InvalidSetup / Exception)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libsphgpu error {code}: {msg}")
        self.code = code


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise FileNotFoundError(
                f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(_LIB_PATH)
        lib.sphgpu_last_error.restype = C.c_char_p
        lib.sphgpu_abi_version.restype = C.c_uint32
        if lib.sphgpu_abi_version() != abi.ABI_VERSION:
            raise RuntimeError("libsphgpu ABI version mismatch")
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise SphGpuError(rc, load_library().sphgpu_last_error().decode(errors="replace"))


def make_lattice(particle_count: int, radius: float, centre=(0.0, 0.0, 0.0), eta: float = 1.3, density: float = 2700.0,
                 centred: bool = True, body_flag: int = 0) -> abi.Lattice:
    lat = abi.Lattice()
    lat.center[0], lat.center[1], lat.center[2] = centre
    lat.radius, lat.particle_count, lat.eta, lat.density = radius, particle_count, eta, density
    lat.flags, lat.body_flag = (abi.LATTICE_CENTER if centred else 0), body_flag
    return lat


def lattice_count(lat: abi.Lattice, device: int = 0) -> int:
    """Number of particles HexagonalPacking yields for the body (sphgpu_lattice_count)."""
    n = C.c_uint32(0)
    _check(load_library().sphgpu_lattice_count(C.c_int(device), C.byref(lat), C.byref(n)))
    return int(n.value)


class Engine:
    """One device context (`sphgpu_ctx`): device-resident particle state + the hot-path kernels."""

    def __init__(self, setup: abi.RunSetup, n_particles: int, capacity: Optional[int] = None, device: int = 0):
        self.lib = load_library()
        self.setup = setup
        self.n = int(n_particles)
        self.capacity = int(capacity if capacity is not None else n_particles)
        self.device = device
        self._ctx = C.c_void_p()
        _check(self.lib.sphgpu_create(C.byref(setup.cfg), setup.materials, C.c_uint32(setup.n_materials),
                                      C.c_uint32(self.n), C.c_uint32(self.capacity), C.c_int(device), C.byref(self._ctx)))
        if setup.cfg.flags & abi.FLAG_XSPH:
            _check(self.lib.sphgpu_set_xsph_epsilon(self._ctx, C.c_double(getattr(setup, "xsph_eps", 1.0))))
        if setup.cfg.flags & abi.FLAG_DELTASPH:
            _check(self.lib.sphgpu_set_deltasph(self._ctx, C.c_double(getattr(setup, "deltasph_delta", 0.01)),
                                                C.c_double(getattr(setup, "deltasph_alpha", 0.01))))
        if setup.cfg.flags & abi.FLAG_STRESS_AV:
            _check(self.lib.sphgpu_set_stress_av(self._ctx, C.c_double(getattr(setup, "stress_av_exponent", 4.0)),
                                                 C.c_double(getattr(setup, "stress_av_factor", 0.04))))

    # -- lifetime ----------------------------------------------------------------------------------------------
    def close(self) -> None:
        if self._ctx:
            self.lib.sphgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- state transfer ----------------------------------------------------------------------------------------
    def upload(self, quantity: str, order: int, array: np.ndarray, first: int = 0) -> None:
        qid, ncomp, dtype = abi.QUANTITIES[quantity]
        a = np.ascontiguousarray(array, dtype=dtype)
        count = a.shape[0]
        assert a.size == count * ncomp, f"{quantity}: expected {ncomp} components"
        _check(self.lib.sphgpu_upload(self._ctx, C.c_int(qid), C.c_int(order), C.c_int(abi.LAYOUT_PACKED),
                                      a.ctypes.data_as(C.c_void_p), C.c_uint32(first), C.c_uint32(count)))

    def download(self, quantity: str, order: int = 0, first: int = 0, count: Optional[int] = None,
                 out: Optional[np.ndarray] = None) -> np.ndarray:
        qid, ncomp, dtype = abi.QUANTITIES[quantity]
        count = self.n - first if count is None else count
        if out is None:
            out = np.empty((count, ncomp) if ncomp > 1 else (count,), dtype=dtype)
        _check(self.lib.sphgpu_download(self._ctx, C.c_int(qid), C.c_int(order), C.c_int(abi.LAYOUT_PACKED),
                                        out.ctypes.data_as(C.c_void_p), C.c_uint32(first), C.c_uint32(count)))
        return out

    def upload_async(self, quantity: str, order: int, array: np.ndarray, first: int = 0) -> None:
        """Queues the upload; `array` (C-contiguous, ideally pinned) must stay alive until transfer_sync()."""
        qid, ncomp, dtype = abi.QUANTITIES[quantity]
        assert array.dtype == dtype and array.flags["C_CONTIGUOUS"]
        count = array.shape[0]
        _check(self.lib.sphgpu_upload_async(self._ctx, C.c_int(qid), C.c_int(order), C.c_int(abi.LAYOUT_PACKED),
                                            array.ctypes.data_as(C.c_void_p), C.c_uint32(first), C.c_uint32(count)))

    def download_async(self, quantity: str, order: int, out: np.ndarray, first: int = 0) -> None:
        """Queues pack + device -> host copy into `out` (valid after transfer_sync())."""
        qid, ncomp, dtype = abi.QUANTITIES[quantity]
        assert out.dtype == dtype and out.flags["C_CONTIGUOUS"]
        _check(self.lib.sphgpu_download_async(self._ctx, C.c_int(qid), C.c_int(order), C.c_int(abi.LAYOUT_PACKED),
                                              out.ctypes.data_as(C.c_void_p), C.c_uint32(first), C.c_uint32(out.shape[0])))

    def download_batch_end(self) -> None:
        _check(self.lib.sphgpu_download_batch_end(self._ctx))

    def transfer_sync(self) -> None:
        _check(self.lib.sphgpu_transfer_sync(self._ctx))

    def upload_device(self, quantity: str, order: int, dev_ptr: int, first: int, count: int) -> None:
        qid, _, _ = abi.QUANTITIES[quantity]
        _check(self.lib.sphgpu_upload_device(self._ctx, C.c_int(qid), C.c_int(order), C.c_void_p(dev_ptr),
                                             C.c_uint32(first), C.c_uint32(count)))

    def download_device(self, quantity: str, order: int, dev_ptr: int, first: int, count: int) -> None:
        qid, _, _ = abi.QUANTITIES[quantity]
        _check(self.lib.sphgpu_download_device(self._ctx, C.c_int(qid), C.c_int(order), C.c_void_p(dev_ptr),
                                               C.c_uint32(first), C.c_uint32(count)))

    def halo_pack(self, first: int, count: int, dev_ptr: int) -> None:
        _check(self.lib.sphgpu_halo_pack(self._ctx, C.c_uint32(first), C.c_uint32(count), C.c_void_p(dev_ptr)))

    def halo_unpack(self, first: int, count: int, dev_ptr: int) -> None:
        _check(self.lib.sphgpu_halo_unpack(self._ctx, C.c_uint32(first), C.c_uint32(count), C.c_void_p(dev_ptr)))

    # -- multi-GPU step inside the library (NCCL) ------------------------------------------------------------------
    def comm_init(self, rank: int, world: int) -> None:
        """Joins the library's own NCCL communicator; the 128-byte id is distributed with torch.distributed."""
        import torch
        import torch.distributed as dist
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            _check(self.lib.sphgpu_comm_unique_id(ident))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        _check(self.lib.sphgpu_comm_init(self._ctx, ident, C.c_int(rank), C.c_int(world)))

    def halo_configure(self, left: int, right: int, send_left: int, send_right: int, recv_left: int, recv_right: int) -> None:
        _check(self.lib.sphgpu_halo_configure(self._ctx, C.c_int(left), C.c_int(right), C.c_uint32(send_left),
                                              C.c_uint32(send_right), C.c_uint32(recv_left), C.c_uint32(recv_right)))

    def peer_connect(self, rank: int, world: int) -> bool:
        """Switches the halo exchange and the time-step reduction to peer memory (CUDA IPC over NVLink): every rank exports
        its handles, the blobs are gathered with torch.distributed, every rank maps its neighbours. Returns False (and stays
        on the NCCL path) if the devices cannot map each other's memory."""
        import torch
        import torch.distributed as dist
        nbytes = 1280  # SPHGPU_PEER_BLOB_BYTES
        blob = (C.c_ubyte * nbytes)()
        ok = self.lib.sphgpu_peer_export(self._ctx, blob) == 0
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor(list(bytes(blob)) + [1 if ok else 0], dtype=torch.uint8, device=dev)
        allb = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        allb = [t.cpu().numpy() for t in allb]
        if not all(int(b[-1]) == 1 for b in allb):
            return False
        packed = b"".join(bytes(b[:-1].tobytes()) for b in allb)
        buf = (C.c_ubyte * len(packed)).from_buffer_copy(packed)
        ok = self.lib.sphgpu_peer_connect(self._ctx, buf, C.c_int(world)) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise SphGpuError(abi.E_STATE, "peer-memory connection failed on some rank: " + self.lib.sphgpu_last_error().decode())
        return True

    def halo_exchange(self) -> None:
        _check(self.lib.sphgpu_halo_exchange(self._ctx))

    def step_pc_mgpu(self, dt: float, max_dt: float, t: float = 0.0) -> Tuple[float, int, abi.Stats]:
        st, ts = abi.Stats(), abi.TimeStep()
        _check(self.lib.sphgpu_step_pc_mgpu(self._ctx, C.c_double(t), C.c_double(dt), C.c_double(max_dt), C.byref(st), C.byref(ts)))
        return float(ts.dt), int(ts.criterion), st

    def upload_state(self, arrays: Dict[str, np.ndarray], names: Optional[Iterable[str]] = None, first: int = 0) -> int:
        """Uploads every snapshot-named array present (pos, vel, rho, ...). Returns the bytes copied."""
        total = 0
        for name in (names if names is not None else abi.SNAPSHOT_FIELDS.keys()):
            if name in arrays and name in abi.SNAPSHOT_FIELDS:
                q, order = abi.SNAPSHOT_FIELDS[name]
                self.upload(q, order, arrays[name], first)
                total += arrays[name].nbytes
        return total

    def download_state(self, names: Iterable[str]) -> Dict[str, np.ndarray]:
        out = {}
        for name in names:
            q, order = abi.SNAPSHOT_FIELDS[name]
            out[name] = self.download(q, order)
        return out

    def set_particle_count(self, n: int) -> None:
        """Changes the number of owned particles (after migration between ranks); the state must be uploaded again."""
        _check(self.lib.sphgpu_set_particle_count(self._ctx, C.c_uint32(n)))
        self.n = n

    def set_active(self, n_active: int) -> None:
        _check(self.lib.sphgpu_set_active(self._ctx, C.c_uint32(n_active)))

    # -- hot path ----------------------------------------------------------------------------------------------
    def integrate(self, t: float = 0.0) -> abi.Stats:
        st = abi.Stats()
        _check(self.lib.sphgpu_integrate(self._ctx, C.c_double(t), C.byref(st)))
        return st

    def set_frozen(self, flags=(), domain=None) -> None:
        """FrozenParticles boundary condition: bodies `flags` and, with domain = (centre, radius, freeze_radius), the particles
        near the surface of that sphere keep zero highest derivatives (sphgpu_set_frozen); no arguments switch it off."""
        f = abi.Frozen()
        for b in flags:
            f.flag_mask |= 1 << int(b)
        if domain is not None:
            centre, radius, freeze_radius = domain
            f.has_domain = 1
            f.center[0], f.center[1], f.center[2] = centre
            f.radius, f.freeze_radius = radius, freeze_radius
        _check(self.lib.sphgpu_set_frozen(self._ctx, C.byref(f) if (f.flag_mask or f.has_domain) else None))

    def lattice_generate(self, lat: abi.Lattice, first: int = 0) -> int:
        """InitialConditions::addMonolithicBody on the device: positions, h, masses, flag of slots [first, first + count)."""
        n = C.c_uint32(0)
        _check(self.lib.sphgpu_lattice_generate(self._ctx, C.byref(lat), C.c_uint32(first), C.byref(n)))
        return int(n.value)

    # -- self-gravity ------------------------------------------------------------------------------------------
    def gravity_configure(self, opening_angle: float = 0.5, order: int = 3, constant: float = abi.GRAVITY_CONSTANT,
                          lut_grad: Optional[np.ndarray] = None, kernel_radius: float = 0.0, leaf_size: int = 0) -> None:
        """Switches device self-gravity on (Factory::getGravity's settings); lut_grad=None selects point particles."""
        cfg = abi.Gravity()
        cfg.opening_angle, cfg.multipole_order, cfg.leaf_size, cfg.constant = opening_angle, order, leaf_size, constant
        self._grav_lut = None
        if lut_grad is not None:
            self._grav_lut = np.ascontiguousarray(lut_grad, dtype=np.float64)
            cfg.lut_grad = self._grav_lut.ctypes.data_as(C.POINTER(C.c_double))
            cfg.lut_entries = len(self._grav_lut) - 1
            cfg.kernel_radius = kernel_radius
        _check(self.lib.sphgpu_gravity_configure(self._ctx, C.byref(cfg)))

    def gravity_off(self) -> None:
        _check(self.lib.sphgpu_gravity_configure(self._ctx, None))

    def gravity_eval(self, accumulate: bool = False) -> abi.GravityStats:
        """IGravity::build + evalSelfGravity on the device state; overwrites (or adds to) the accelerations."""
        st = abi.GravityStats()
        _check(self.lib.sphgpu_gravity_eval(self._ctx, C.c_int(1 if accumulate else 0), C.byref(st)))
        return st

    def gravity_last_stats(self) -> abi.GravityStats:
        st = abi.GravityStats()
        _check(self.lib.sphgpu_gravity_last_stats(self._ctx, C.byref(st)))
        return st

    def predict(self, dt: float) -> None:
        _check(self.lib.sphgpu_step_predict(self._ctx, C.c_double(dt)))

    def correct(self, dt: float) -> None:
        _check(self.lib.sphgpu_step_correct(self._ctx, C.c_double(dt)))

    def euler(self, dt: float) -> None:
        _check(self.lib.sphgpu_step_euler(self._ctx, C.c_double(dt)))

    def compute_timestep(self, max_dt: float) -> Tuple[float, int]:
        ts = abi.TimeStep()
        _check(self.lib.sphgpu_compute_timestep(self._ctx, C.c_double(max_dt), C.byref(ts)))
        return float(ts.dt), int(ts.criterion)

    def set_last_timestep(self, dt: float) -> None:
        _check(self.lib.sphgpu_set_last_timestep(self._ctx, C.c_double(dt)))

    def step_pc(self, dt: float, max_dt: float, t: float = 0.0) -> Tuple[float, int, abi.Stats]:
        st, ts = abi.Stats(), abi.TimeStep()
        _check(self.lib.sphgpu_step_pc(self._ctx, C.c_double(t), C.c_double(dt), C.c_double(max_dt), C.byref(st), C.byref(ts)))
        return float(ts.dt), int(ts.criterion), st

    # -- inspection --------------------------------------------------------------------------------------------
    def neighbours(self) -> Tuple[np.ndarray, np.ndarray]:
        off = np.zeros(self.n + 1, np.uint64)
        _check(self.lib.sphgpu_neighbour_dump(self._ctx, off.ctypes.data_as(C.POINTER(C.c_uint64)), None, C.c_uint64(0)))
        idx = np.zeros(int(off[-1]), np.uint32)
        _check(self.lib.sphgpu_neighbour_dump(self._ctx, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(len(idx))))
        return off, idx

    def find_components(self, radius: float, separate_by_flag: bool = False) -> Tuple[np.ndarray, int, int]:
        """Post::findComponents of the reference on the particles in the context (core/post/Analysis.cpp:36-75,115-128): returns
        (component index of every particle, number of components, label-propagation sweeps the device needed)."""
        idx = np.zeros(max(self.n, 1), np.uint32)
        count, sweeps = C.c_uint32(0), C.c_uint32(0)
        _check(self.lib.sphgpu_find_components(self._ctx, C.c_double(radius), C.c_uint32(1 if separate_by_flag else 0),
                                               idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(count), C.byref(sweeps)))
        return idx[:self.n], int(count.value), int(sweeps.value)

    def last_timings(self) -> np.ndarray:
        ms = np.zeros(4)
        _check(self.lib.sphgpu_last_timings(self._ctx, ms.ctypes.data_as(C.POINTER(C.c_double))))
        return ms

    def run_pc(self, steps: int, dt: float, max_dt: float):
        """`steps` PredictorCorrector steps with one host synchronisation; returns (dt chosen after every step, criterion ids,
        stats of the last step)."""
        st = abi.Stats()
        hist = (abi.TimeStep * max(steps, 1))()
        _check(self.lib.sphgpu_run_pc(self._ctx, C.c_uint32(steps), C.c_double(dt), C.c_double(max_dt), C.byref(st), hist))
        return (np.array([hist[s].dt for s in range(steps)]), np.array([hist[s].criterion for s in range(steps)], np.uint32), st)

    def last_pair_timings(self) -> np.ndarray:
        """Device ms of {unit preparation, candidate lists, pair sums} of the last pair stage."""
        out = (C.c_double * 3)()
        _check(self.lib.sphgpu_last_pair_timings(self._ctx, out))
        return np.array(list(out))

    def measure_fp64_peak(self) -> float:
        """FP64 fused multiply-adds per second of this device (DFMA microbenchmark inside the library)."""
        v = C.c_double(0.0)
        _check(self.lib.sphgpu_measure_fp64_peak(self._ctx, C.byref(v)))
        return float(v.value)

    def last_halo_ms(self) -> float:
        ms = C.c_double(0.0)
        _check(self.lib.sphgpu_last_halo_ms(self._ctx, C.byref(ms)))
        return float(ms.value)

    def set_variant(self, variant: int) -> None:
        _check(self.lib.sphgpu_set_variant(self._ctx, C.c_int(variant)))

    def halo_set_guard(self, axis: int, lo_plane, hi_plane) -> None:
        """Cut planes of this rank's domain (None: no neighbour on that side); see sphgpu_halo_set_guard."""
        _check(self.lib.sphgpu_halo_set_guard(self._ctx, C.c_int(axis), C.c_double(0.0 if lo_plane is None else lo_plane),
                                              C.c_double(0.0 if hi_plane is None else hi_plane), C.c_int(lo_plane is not None),
                                              C.c_int(hi_plane is not None)))

    def halo_margin(self) -> float:
        m = C.c_double(0.0)
        _check(self.lib.sphgpu_halo_margin(self._ctx, C.byref(m)))
        return float(m.value)

    def set_list_skin(self, skin: float) -> None:
        """Relative enlargement of the candidate lists' search radius (0: the lists are rebuilt in every integrate)."""
        _check(self.lib.sphgpu_set_list_skin(self._ctx, C.c_double(skin)))

    def list_stats(self):
        """(builds so far, calls served by the current lists, displacement metric of the last call)."""
        r, a, m = C.c_uint32(0), C.c_uint32(0), C.c_double(0.0)
        _check(self.lib.sphgpu_list_stats(self._ctx, C.byref(r), C.byref(a), C.byref(m)))
        return int(r.value), int(a.value), float(m.value)

    def set_stream(self, cuda_stream: int) -> None:
        _check(self.lib.sphgpu_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def use_private_stream(self) -> None:
        _check(self.lib.sphgpu_use_private_stream(self._ctx))

    def synchronize(self) -> None:
        _check(self.lib.sphgpu_synchronize(self._ctx))
