"""Multi-GPU domain decomposition for the asymmetric SPH evaluation: one process per GPU, particles sharded by
position, ghost (halo) particles exchanged once per derivative evaluation.

The reference has no distributed path (shared-memory parallelFor only). The asymmetric formulation writes
only to particle i and reads neighbour inputs, so one exchange of the neighbour inputs per integrate() suffices
(SURVEY 8e); ghosts are appended after the owned particles -- the analogue of GhostParticles on the CPU
(core/sph/boundary/Boundary.h:73-) -- and are never targets.

Round-1 decomposition: slabs of equal particle counts along z, re-cut and migrated on demand (`repartition`) (a 1-D space-filling curve; z because the cell rows of
the pair kernel run along x and must stay long). Each rank keeps its
slots ordered [left band | interior | right band], so the particles a neighbour needs are two contiguous slot ranges
and packing is a plain range copy (sphgpu_download_device). The dynamic neighbour inputs (r,h | v | rho | u | S | D) are
exchanged with NCCL point-to-point every step; static ones (m, flag, material) once.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi, workloads

# (snapshot name, components) of the neighbour inputs
DYNAMIC_FIELDS: Tuple[Tuple[str, int], ...] = (("pos", 4), ("vel", 4), ("rho", 1), ("u", 1), ("S", 5), ("damage", 1))
STATIC_FIELDS: Tuple[Tuple[str, int], ...] = (("mass", 1),)
STATIC_U32: Tuple[str, ...] = ("flag",)
HALO_MARGIN = 0.25  # relative head-room of the send bands beyond the kernel reach R h_max


def sphere_cut_planes(radius: float, parts: int) -> np.ndarray:
    """x positions cutting a sphere into `parts` slabs of equal volume (parts+1 values, -R .. R)."""
    def frac(x):  # volume fraction of the cap [-R, x]
        return (3.0 * radius * radius * (x + radius) - (x ** 3 + radius ** 3)) / (4.0 * radius ** 3)
    cuts = [-radius]
    for k in range(1, parts):
        lo, hi = -radius, radius
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if frac(mid) < k / parts:
                lo = mid
            else:
                hi = mid
        cuts.append(0.5 * (lo + hi))
    cuts.append(radius)
    return np.array(cuts)


def band_partition(x: np.ndarray, lo_plane: Optional[float], hi_plane: Optional[float], width: float):
    """Stable permutation ordering particles [left band | interior | right band] and the two band sizes.

    Left band = particles within `width` of the lower cut plane (needed by the left neighbour), right band likewise.
    A missing neighbour (plane None) gives an empty band."""
    left = (x < lo_plane + width) if lo_plane is not None else np.zeros(len(x), bool)
    right = (x >= hi_plane - width) if hi_plane is not None else np.zeros(len(x), bool)
    if np.any(left & right):
        raise ValueError("slab thinner than the halo width: use fewer ranks or more particles")
    group = np.where(left, 0, np.where(right, 2, 1))
    perm = np.argsort(group, kind="stable")
    return perm, int(left.sum()), int(right.sum())


class SlabDomain:
    """Geometry of the basalt-sphere workload cut into x-slabs; every rank generates only its own lattice points."""

    def __init__(self, n_target: int, world: int, rank: int, radius: float = 5.0e4, solid: bool = True, axis: int = 2):
        self.n_target, self.world, self.rank, self.radius, self.solid = n_target, world, rank, radius, solid
        self.axis = axis  # slabs are cut perpendicular to this axis
        self.cuts = sphere_cut_planes(radius, world)
        volume = 4.0 / 3.0 * math.pi * radius ** 3
        self.h = (volume / n_target) ** (1.0 / 3.0) * workloads.BASALT["eta"]
        # kernel radius * h_max with head-room for drifting particles and growing h: the bands are fixed slot ranges, the
        # library's halo guard (sphgpu_halo_set_guard) reports how much of the head-room is left and fails hard at zero
        self.halo_width = 2.0 * self.h * (1.0 + HALO_MARGIN)
        self.lo_plane = float(self.cuts[rank]) if rank > 0 else None
        self.hi_plane = float(self.cuts[rank + 1]) if rank < world - 1 else None
        self.n_left = self.n_right = 0
        self._n_total: Optional[int] = None

    def generate_owned(self) -> Dict[str, np.ndarray]:
        lo = self.lo_plane if self.lo_plane is not None else -2.0 * self.radius
        hi = self.hi_plane if self.hi_plane is not None else 2.0 * self.radius
        x_range = None if self.world == 1 else (lo, hi)
        pos, _ = workloads.hexagonal_sphere(self.n_target, self.radius, x_range=x_range, axis=self.axis)
        n_total = self.total_particles(len(pos))
        state = workloads.basalt_sphere_state(self.n_target, self.radius, self.solid, x_range=x_range, total_hint=n_total,
                                              axis=self.axis)
        if self.world > 1:
            perm, self.n_left, self.n_right = band_partition(state["pos"][:, self.axis], self.lo_plane, self.hi_plane,
                                                             self.halo_width)
            state = {k: (v[perm] if isinstance(v, np.ndarray) and v.shape[:1] == (len(perm),) else v) for k, v in state.items()}
        return state

    def adopt_cuts(self, cuts: np.ndarray, halo_width: float) -> None:
        """New cut planes (from balanced_cut_planes) and halo width after a repartition."""
        self.cuts = np.asarray(cuts, dtype=np.float64)
        self.lo_plane = float(self.cuts[self.rank]) if self.rank > 0 else None
        self.hi_plane = float(self.cuts[self.rank + 1]) if self.rank < self.world - 1 else None
        self.halo_width = float(halo_width)
        self._n_total = None

    def total_particles(self, n_owned: int) -> int:
        if self._n_total is None:
            if self.world == 1:
                self._n_total = n_owned
            else:
                import torch
                import torch.distributed as dist
                t = torch.tensor([n_owned], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
                dist.all_reduce(t)
                self._n_total = int(t.item())
        return self._n_total

    def capacity(self, n_owned: int) -> int:
        if self.world == 1:
            return n_owned
        density = self.n_target * 1.07 / (4.0 / 3.0 * math.pi * self.radius ** 3)
        per_face = math.pi * self.radius ** 2 * self.halo_width * density
        return n_owned + int(2.2 * per_face) + 1024


# everything a particle carries when it changes rank: state, the derivatives the next predictor step extrapolates with,
# and the per-particle material constants
MIGRATE_FIELDS: Tuple[str, ...] = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "S", "dS", "damage", "ddamage", "reduce",
                                   "eps_min", "m_zero", "growth", "n_flaws", "flag")


def _dist_device(dist) -> str:
    return "cuda" if dist.get_backend() == "nccl" else "cpu"


def balanced_cut_planes(coord: np.ndarray, world: int, dist=None, bins: int = 4096) -> np.ndarray:
    """Cut planes (world + 1 ascending values) splitting the particles of ALL ranks into `world` slabs of (nearly) equal
    count along one coordinate: global histogram (all_reduce), cuts at the count quantiles, linear inside a bin.
    The analogue of re-cutting a space-filling curve into equal pieces; works for any particle distribution."""
    import torch
    lo = float(coord.min()) if len(coord) else np.inf
    hi = float(coord.max()) if len(coord) else -np.inf
    if dist is not None and world > 1:
        t = torch.tensor([lo, -hi], dtype=torch.float64, device=_dist_device(dist))
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        lo, hi = float(t[0].item()), -float(t[1].item())
    span = max(hi - lo, 1e-300)
    edges = lo + span * np.arange(bins + 1) / bins
    edges[-1] = hi + 1e-9 * span
    hist = np.histogram(coord, edges)[0].astype(np.float64)
    if dist is not None and world > 1:
        t = torch.from_numpy(hist).to(_dist_device(dist))
        dist.all_reduce(t)
        hist = t.cpu().numpy()
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    total = cum[-1]
    cuts = [lo - 1e-6 * span]
    for k in range(1, world):
        target = total * k / world
        b = int(np.searchsorted(cum, target, side="right")) - 1
        b = min(max(b, 0), bins - 1)
        frac = (target - cum[b]) / hist[b] if hist[b] > 0 else 0.0
        cuts.append(float(edges[b] + frac * (edges[b + 1] - edges[b])))
    cuts.append(hi + 1e-6 * span)
    return np.array(cuts)


def migrate(state: Dict[str, np.ndarray], dest: np.ndarray, world: int, rank: int, dist) -> Dict[str, np.ndarray]:
    """Moves every particle to rank dest[i]: one message per pair of ranks (all per-particle arrays of `state` packed as
    rows of doubles; u32 fields are exact in a double). Returns the new owned state: kept particles first (original
    order), then the received ones by source rank."""
    import torch
    n = len(dest)
    names = [k for k, v in state.items() if isinstance(v, np.ndarray) and v.shape[:1] == (n,)]
    widths = [int(np.prod(state[k].shape[1:], dtype=np.int64)) for k in names]
    width = sum(widths)
    rows = np.empty((n, width), np.float64)
    off = 0
    for k, w in zip(names, widths):
        rows[:, off:off + w] = state[k].reshape(n, w)
        off += w
    dev = _dist_device(dist)
    mine = torch.tensor([int((dest == r).sum()) for r in range(world)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(counts, mine)
    counts = torch.stack(counts).cpu().numpy()  # counts[src, dst]
    ops, recv = [], {}
    send_bufs = []
    for peer in range(world):
        if peer == rank:
            continue
        if counts[rank, peer] > 0:
            buf = torch.from_numpy(np.ascontiguousarray(rows[dest == peer])).to(dev)
            send_bufs.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, peer))
        if counts[peer, rank] > 0:
            recv[peer] = torch.empty((int(counts[peer, rank]), width), dtype=torch.float64, device=dev)
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    parts = [rows[dest == rank]] + [recv[p].cpu().numpy() for p in sorted(recv)]
    rows = np.concatenate(parts, axis=0) if parts else rows[:0]
    out = {k: v for k, v in state.items() if k not in names}
    off = 0
    for k, w in zip(names, widths):
        block = rows[:, off:off + w]
        out[k] = np.ascontiguousarray(block.reshape((len(rows),) + state[k].shape[1:]).astype(state[k].dtype))
        off += w
    return out


def _repartition_device(dom: "SlabDomain", eng, names: Sequence[str], kernel_radius: float):
    """repartition() with the particle data staying in device memory: the rows are packed from the slot planes into
    CUDA tensors (sphgpu_download_device), the new owner of every particle, the rows to send and the
    [lower band | interior | upper band] order are computed with torch on the GPU, the rows travel over NCCL
    (batch_isend_irecv of CUDA tensors: NVLink), and the new owned state is written back with sphgpu_upload_device.
    Only a 4096-bin histogram and a few counts cross PCIe."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    n = eng.n
    fields = [(k,) + abi.SNAPSHOT_FIELDS[k] for k in names] + [("material_id", "MATERIAL_ID", 0)]
    cols: Dict[str, "torch.Tensor"] = {}
    for key, q, order in fields:
        _, ncomp, dtype = abi.QUANTITIES[q]
        t = torch.empty((max(n, 1), ncomp), dtype=torch.float64 if dtype == np.float64 else torch.int32, device=dev)
        if n:
            eng.download_device(q, order, t.data_ptr(), 0, n)
        cols[key] = t[:n]
    eng.synchronize()  # the packs ran on the engine's stream
    axis = dom.axis
    z = cols["pos"][:, axis].contiguous()
    # global histogram of the cut coordinate -> cut planes at the count quantiles (as balanced_cut_planes)
    bins = 4096
    lohi = torch.stack([z.min() if n else torch.tensor(float("inf"), device=dev, dtype=torch.float64),
                        -(z.max()) if n else torch.tensor(float("inf"), device=dev, dtype=torch.float64)])
    dist.all_reduce(lohi, op=dist.ReduceOp.MIN)
    lo, hi = float(lohi[0].item()), -float(lohi[1].item())
    span = max(hi - lo, 1e-300)
    idx = torch.clamp(((z - lo) * (bins / span)).to(torch.int64), 0, bins - 1)
    hist = torch.bincount(idx, minlength=bins).to(torch.float64)
    dist.all_reduce(hist)
    hist = hist.cpu().numpy()
    edges = lo + span * np.arange(bins + 1) / bins
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    cuts = [lo - 1e-6 * span]
    for k in range(1, dom.world):
        target = cum[-1] * k / dom.world
        b = min(max(int(np.searchsorted(cum, target, side="right")) - 1, 0), bins - 1)
        frac = (target - cum[b]) / hist[b] if hist[b] > 0 else 0.0
        cuts.append(float(edges[b] + frac * (edges[b + 1] - edges[b])))
    cuts.append(hi + 1e-6 * span)
    cuts = np.array(cuts)
    dest = torch.clamp(torch.bucketize(z, torch.tensor(cuts[1:-1], dtype=torch.float64, device=dev), right=True), 0, dom.world - 1)
    # rows of doubles (u32 fields are exact in a double), one message per pair of ranks
    widths = [cols[k].shape[1] for k, _, _ in fields]
    rows = torch.cat([cols[k].to(torch.float64) for k, _, _ in fields], dim=1) if n else torch.empty((0, sum(widths)), dtype=torch.float64, device=dev)
    mine = torch.bincount(dest, minlength=dom.world).to(torch.int64)
    counts = [torch.zeros_like(mine) for _ in range(dom.world)]
    dist.all_gather(counts, mine)
    counts = torch.stack(counts).cpu().numpy()  # counts[src, dst]
    ops, recv, keep_alive = [], {}, []
    for peer in range(dom.world):
        if peer == dom.rank:
            continue
        if counts[dom.rank, peer] > 0:
            buf = rows[dest == peer].contiguous()
            keep_alive.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, peer))
        if counts[peer, dom.rank] > 0:
            recv[peer] = torch.empty((int(counts[peer, dom.rank]), rows.shape[1]), dtype=torch.float64, device=dev)
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    rows = torch.cat([rows[dest == dom.rank]] + [recv[p] for p in sorted(recv)], dim=0)
    # halo width from the CURRENT largest smoothing length of the whole run, then the band order
    hmax = rows[:, 3].max().reshape(1) if len(rows) else torch.zeros(1, dtype=torch.float64, device=dev)
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dom.adopt_cuts(cuts, kernel_radius * float(hmax.item()) * (1.0 + HALO_MARGIN))
    x = rows[:, axis]
    left = (x < dom.lo_plane + dom.halo_width) if dom.lo_plane is not None else torch.zeros(len(rows), dtype=torch.bool, device=dev)
    right = (x >= dom.hi_plane - dom.halo_width) if dom.hi_plane is not None else torch.zeros(len(rows), dtype=torch.bool, device=dev)
    if bool((left & right).any().item()):
        raise ValueError("slab thinner than the halo width: use fewer ranks or more particles")
    group = torch.where(left, 0, torch.where(right, 2, 1))
    perm = torch.sort(group, stable=True).indices
    rows = rows[perm]
    dom.n_left, dom.n_right = int(left.sum().item()), int(right.sum().item())
    n_new = len(rows)
    if n_new > eng.capacity:
        raise ValueError(f"rank {dom.rank}: {n_new} particles after migration exceed the engine capacity {eng.capacity}")
    eng.set_particle_count(n_new)
    off, keep = 0, []
    for (key, q, order), w in zip(fields, widths):
        _, ncomp, dtype = abi.QUANTITIES[q]
        block = rows[:, off:off + w]
        block = block.contiguous() if dtype == np.float64 else block.to(torch.int32).contiguous()
        keep.append(block)
        off += w
    torch.cuda.current_stream().synchronize()  # the tensors are complete before the engine's stream reads them
    for (key, q, order), block in zip(fields, keep):
        if n_new:
            eng.upload_device(q, order, block.data_ptr(), 0, n_new)
    eng.synchronize()
    halo = HaloExchange(dom, eng, None, names=tuple(names))
    return None, halo


def repartition(dom: "SlabDomain", eng, names: Sequence[str] = MIGRATE_FIELDS, adapter=None, kernel_radius: float = 2.0):
    """Re-cuts the slabs to equal particle counts and migrates the particles that changed slab ('re-cut every
    m steps and migrate particles'): restores the [lower band | interior | upper band] slot order and rebuilds the halo
    exchange. With NCCL and a device engine the particle rows never leave device memory (_repartition_device; the returned
    state is None); otherwise (gloo / host engines: the CPU tests) the owned state is downloaded, exchanged point-to-point
    and uploaded again. Returns (new owned state on the host or None, new HaloExchange)."""
    import torch
    import torch.distributed as dist
    if adapter is None and dist.get_backend() == "nccl" and hasattr(eng, "download_device") and hasattr(eng, "synchronize"):
        return _repartition_device(dom, eng, names, kernel_radius)
    state = eng.download_state([k for k in names])
    state["material_id"] = eng.download("MATERIAL_ID", 0, 0, eng.n) if hasattr(eng, "download") else np.zeros(len(state["pos"]), np.uint32)
    axis = dom.axis
    cuts = balanced_cut_planes(state["pos"][:, axis], dom.world, dist)
    dest = np.clip(np.searchsorted(cuts[1:-1], state["pos"][:, axis], side="right"), 0, dom.world - 1)
    state = migrate(state, dest, dom.world, dom.rank, dist)
    # halo width from the CURRENT largest smoothing length of the whole run
    hmax = torch.tensor([float(state["pos"][:, 3].max()) if len(state["pos"]) else 0.0], dtype=torch.float64, device=_dist_device(dist))
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dom.adopt_cuts(cuts, kernel_radius * float(hmax.item()) * (1.0 + HALO_MARGIN))
    perm, dom.n_left, dom.n_right = band_partition(state["pos"][:, axis], dom.lo_plane, dom.hi_plane, dom.halo_width)
    state = {k: (v[perm] if isinstance(v, np.ndarray) and v.shape[:1] == (len(perm),) else v) for k, v in state.items()}
    n_new = len(perm)
    if n_new > eng.capacity:
        raise ValueError(f"rank {dom.rank}: {n_new} particles after migration exceed the engine capacity {eng.capacity}")
    eng.set_particle_count(n_new)
    eng.upload_state(state, names)
    if hasattr(eng, "upload") and n_new:
        eng.upload("MATERIAL_ID", 0, state["material_id"].astype(np.uint32))
    halo = HaloExchange(dom, eng, state, adapter=adapter)
    return state, halo


class EngineAdapter:
    """Packs / unpacks slot ranges of a device engine into flat float64 torch tensors on the GPU."""

    def __init__(self, eng):
        self.eng = eng

    def new_buffer(self, doubles: int):
        import torch
        return torch.empty(max(doubles, 1), dtype=torch.float64, device="cuda")

    def pack(self, fields: Sequence[Tuple[str, int]], first: int, count: int, buf) -> None:
        if tuple(fields) == DYNAMIC_FIELDS:  # one fused kernel for the per-step exchange
            self.eng.halo_pack(first, count, buf.data_ptr())
            return
        off = 0
        for name, ncomp in fields:
            q, order = abi.SNAPSHOT_FIELDS[name]
            self.eng.download_device(q, order, buf.data_ptr() + 8 * off, first, count)
            off += ncomp * count

    def unpack(self, fields: Sequence[Tuple[str, int]], first: int, count: int, buf) -> None:
        if tuple(fields) == DYNAMIC_FIELDS:
            self.eng.halo_unpack(first, count, buf.data_ptr())
            return
        off = 0
        for name, ncomp in fields:
            q, order = abi.SNAPSHOT_FIELDS[name]
            self.eng.upload_device(q, order, buf.data_ptr() + 8 * off, first, count)
            off += ncomp * count


class HaloExchange:
    """Ghost-layer exchange between slab neighbours (rank-1, rank+1) with torch.distributed point-to-point."""

    def __init__(self, dom: SlabDomain, eng, state: Optional[Dict[str, np.ndarray]], adapter=None, fields=DYNAMIC_FIELDS, names=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.dom, self.eng = dom, eng
        self.adapter = adapter or EngineAdapter(eng)
        # `state` is only consulted for the particle count and for which quantities exist; a device-side repartition
        # passes None and the names it migrated
        have = set(state.keys()) if state is not None else set(names or ())
        self.n = len(state["mass"]) if state is not None else eng.n
        self.fields = tuple((k, c) for k, c in fields if k in have)
        self.width = sum(c for _, c in self.fields)
        self.left, self.right = (dom.rank - 1 if dom.rank > 0 else None), (dom.rank + 1 if dom.rank < dom.world - 1 else None)
        # how many ghosts arrive from each side: the neighbour's band facing us
        counts = self._exchange_counts(dom.n_left, dom.n_right)
        self.g_left, self.g_right = counts
        self.ghost_left_first = self.n
        self.ghost_right_first = self.n + self.g_left
        self.n_active = self.n + self.g_left + self.g_right
        if self.n_active > eng.capacity:
            raise ValueError(f"ghost capacity too small: need {self.n_active}, have {eng.capacity}")
        self.send_l = self.adapter.new_buffer(self.width * dom.n_left)
        self.send_r = self.adapter.new_buffer(self.width * dom.n_right)
        self.recv_l = self.adapter.new_buffer(self.width * self.g_left)
        self.recv_r = self.adapter.new_buffer(self.width * self.g_right)
        self.bytes_per_exchange = 8 * self.width * (dom.n_left + dom.n_right)
        self._static(have)
        eng.set_active(self.n_active)
        # the per-step exchange runs inside the library (NCCL on the engine's stream) when the engine supports it
        self.native = hasattr(eng, "comm_init") and dist.get_backend() == "nccl" and self.fields == DYNAMIC_FIELDS
        if self.native:
            if not getattr(eng, "_comm_ready", False):  # the communicator survives a repartition
                eng.comm_init(dom.rank, dom.world)
                eng._comm_ready = True
            eng.halo_configure(-1 if self.left is None else self.left, -1 if self.right is None else self.right,
                               dom.n_left, dom.n_right, self.g_left, self.g_right)
            if hasattr(eng, "halo_set_guard"):
                eng.halo_set_guard(dom.axis, dom.lo_plane, dom.hi_plane)
            # ranks of one node: band particles go straight into the neighbours' ghost slots over NVLink
            import os
            self.transport = "nccl point-to-point"
            if hasattr(eng, "peer_connect") and os.environ.get("SPHGPU_HALO_TRANSPORT", "peer") == "peer":
                if eng.peer_connect(dom.rank, dom.world):
                    self.transport = "peer-memory push kernel (CUDA IPC over NVLink)"

    def _exchange_counts(self, n_left: int, n_right: int) -> Tuple[int, int]:
        torch, dist = self.torch, self.dist
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor([n_left, n_right], dtype=torch.int64, device=dev)
        allc = [torch.zeros_like(mine) for _ in range(self.dom.world)]
        dist.all_gather(allc, mine)
        g_left = int(allc[self.left][1].item()) if self.left is not None else 0
        g_right = int(allc[self.right][0].item()) if self.right is not None else 0
        return g_left, g_right

    def _p2p(self, send_l, send_r, recv_l, recv_r) -> None:
        dist = self.dist
        ops = []
        if self.left is not None:
            if send_l.numel() and self.dom.n_left:
                ops.append(dist.P2POp(dist.isend, send_l, self.left))
            if self.g_left:
                ops.append(dist.P2POp(dist.irecv, recv_l, self.left))
        if self.right is not None:
            if send_r.numel() and self.dom.n_right:
                ops.append(dist.P2POp(dist.isend, send_r, self.right))
            if self.g_right:
                ops.append(dist.P2POp(dist.irecv, recv_r, self.right))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _run(self, fields) -> None:
        width = sum(c for _, c in fields)
        nl, nr = self.dom.n_left, self.dom.n_right
        sl, sr = self.send_l[: width * nl], self.send_r[: width * nr]
        rl, rr = self.recv_l[: width * self.g_left], self.recv_r[: width * self.g_right]
        if nl:
            self.adapter.pack(fields, 0, nl, sl)
        if nr:
            self.adapter.pack(fields, self.n - nr, nr, sr)
        self._p2p(sl, sr, rl, rr)
        if self.g_left:
            self.adapter.unpack(fields, self.ghost_left_first, self.g_left, rl)
        if self.g_right:
            self.adapter.unpack(fields, self.ghost_right_first, self.g_right, rr)

    def _static(self, have) -> None:
        fields = tuple((k, c) for k, c in STATIC_FIELDS if k in have)
        if fields:
            self._run(fields)
        # Body flag and material id of the ghosts come from their owners (a ghost of another body must not pass the
        # SUM_ONLY_UNDAMAGED group test, and its EoS / rheology constants are its material's): the band's values travel
        # once, as doubles (u32 is exact in a double), over the same point-to-point pattern.
        if not hasattr(self.eng, "upload"):
            return
        torch = self.torch
        dev = _dist_device(self.dist)
        nl, nr = self.dom.n_left, self.dom.n_right
        for q in ("FLAG", "MATERIAL_ID"):
            lo_band = self.eng.download(q, 0, 0, nl).astype(np.float64) if nl else np.zeros(0)
            hi_band = self.eng.download(q, 0, self.n - nr, nr).astype(np.float64) if nr else np.zeros(0)
            sl = torch.from_numpy(np.ascontiguousarray(lo_band)).to(dev)
            sr = torch.from_numpy(np.ascontiguousarray(hi_band)).to(dev)
            rl = torch.zeros(max(self.g_left, 1), dtype=torch.float64, device=dev)[: self.g_left]
            rr = torch.zeros(max(self.g_right, 1), dtype=torch.float64, device=dev)[: self.g_right]
            self._p2p(sl, sr, rl, rr)
            ghosts = np.concatenate([rl.cpu().numpy(), rr.cpu().numpy()]).astype(np.uint32)
            if len(ghosts):
                self.eng.upload(q, 0, ghosts, first=self.n)

    def exchange(self) -> None:
        """Refreshes the dynamic neighbour inputs of all ghosts; call after predict and before integrate."""
        if self.native:
            self.eng.halo_exchange()
        else:
            self._run(self.fields)
