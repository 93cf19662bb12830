"""Synthetic basalt-sphere workloads of BASELINE.json's configs, generated without the reference (bench.py must not
touch oracle/ for its own arm): hexagonal close-packed lattice in a sphere, basalt material defaults, smooth seeded
fields so that every term of the collision preset does non-trivial work.

Reference behaviour mirrored here (paths relative to the reference root):
  lattice           HexagonalPacking::generate, core/sph/initial/Distribution.cpp:126-200 (dx = 1.1 (V/n)^(1/3),
                    dy = sqrt(3)/2 dx, dz = sqrt(6)/3 dx, z-outer/x-inner raster order; without the CENTER shift)
  h, masses         InitialConditions::setQuantities, core/sph/initial/Initial.cpp:308-333 (h *= eta; m ~ h^3, sum = rho0 V)
  basalt defaults   core/system/Settings.cpp:790-930
  run settings      SphJob::getDefaultSettings, core/run/jobs/SimulationJobs.cpp:195-232 (collision preset, gravity off)
  flaws             ScalarGradyKippModel::setFlaws (sampled variant), core/physics/Damage.cpp:17-125
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from . import abi

INF = 1.7976931348623157e308

BASALT = dict(
    til_u0=4.87e8, til_uiv=4.72e6, til_ucv=1.82e7, til_a=0.5, til_b=1.5, rho0=2700.0, til_A=2.67e10, til_B=2.67e10,
    til_alpha=5.0, til_beta=5.0, gamma=1.4, shear_modulus=2.27e10, elasticity_limit=3.5e9, melt_energy=3.4e6,
    weibull_k=4.0e35, weibull_m=9.0, rayleigh=0.4, eta=1.3,
    rho_range=(50.0, INF), rho_small=100.0, u_range=(0.0, INF), u_small=1.0, d_range=(0.0, 1.0), d_small=0.03,
    s_small=1.0e5,
)


def cubic_spline_lut(entries: int = 40000, radius: float = 2.0) -> Tuple[np.ndarray, np.ndarray]:
    """LutKernel<3>(CubicSpline<3>) tables: W and (dW/dq)/q sampled at q^2 = i * R^2 / entries (Kernel.h:85-101,149-188)."""
    q = np.sqrt(np.arange(entries + 1, dtype=np.float64) / (entries * (1.0 / (radius * radius))))
    norm = 1.0 / math.pi
    val = np.where(q < 1, norm * (0.25 * (2 - q) ** 3 - (1 - q) ** 3), np.where(q < 2, norm * 0.25 * (2 - q) ** 3, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        g1 = (1.0 / q) * norm * (-0.75 * (2 - q) ** 2 + 3 * (1 - q) ** 2)
        g2 = (1.0 / q) * norm * (-0.75 * (2 - q) ** 2)
    grad = np.where(q == 0, -3 * norm, np.where(q < 1, g1, np.where(q < 2, g2, 0.0)))
    return grad, val


def hexagonal_sphere(n_target: int, radius: float, centre=(0.0, 0.0, 0.0), x_range=None, axis: int = 0) -> Tuple[np.ndarray, float]:
    """Positions of the hexagonal lattice filling a sphere (about 1.06 n_target particles) and the lattice h.

    `x_range` = (lo, hi) keeps only lattice points with lo <= r[axis] < hi (used by ranks to generate just their slab)."""
    volume = 4.0 / 3.0 * math.pi * radius ** 3
    h = 1.0 / (n_target / volume) ** (1.0 / 3.0)
    dx = 1.1 * h
    dy = math.sqrt(3.0) * 0.5 * dx
    dz = math.sqrt(6.0) / 3.0 * dx
    lo = np.array([-radius + 0.5 * dx, -radius + 0.5 * dy, -radius + 0.5 * dz])
    nx = int(math.floor((radius - lo[0]) / dx)) + 1
    ny = int(math.floor((radius - lo[1]) / dy)) + 1
    nz = int(math.floor((radius - lo[2]) / dz)) + 1
    xs = lo[0] + dx * np.arange(nx)
    out = []
    r2 = radius * radius
    for k in range(nz):  # z outer, y, x inner: the reference's raster order
        z = lo[2] + dz * k
        j = np.arange(ny)
        y = lo[1] + dy * j + (math.sqrt(3.0) / 6.0 * dx if k % 2 == 1 else 0.0)
        shift = np.where((j % 2 == 1) if k % 2 == 0 else (j % 2 == 0), 0.5 * dx, 0.0)
        X = xs[None, :] + shift[:, None]
        Y = np.broadcast_to(y[:, None], X.shape)
        mask = X * X + Y * Y + z * z <= r2
        if x_range is not None:
            if axis == 2:
                if not (x_range[0] <= z < x_range[1]):
                    continue
            else:
                coord = X if axis == 0 else Y
                mask &= (coord >= x_range[0]) & (coord < x_range[1])
        if mask.any():
            out.append(np.stack([X[mask], Y[mask], np.full(int(mask.sum()), z)], axis=1))
    pos = np.concatenate(out, axis=0) if out else np.zeros((0, 3))
    pos += np.asarray(centre)[None, :]
    return pos, h


def _hash01(pos3: np.ndarray, scale: float, seed: int, stream: int) -> np.ndarray:
    """Uniform(0,1) numbers keyed by the (quantised) particle position, so that every rank of a decomposed run draws
    the same per-particle constants as a single-domain run (splitmix64 of the lattice coordinates)."""
    q = np.round(pos3 / scale).astype(np.int64).astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (q[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (q[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) ^ (
            q[:, 2] * np.uint64(0x165667B19E3779F9)) ^ np.uint64((seed * 0x632BE59BD9B4E019 + stream * 0x9E3779B97F4A7C15) % 2 ** 64)
        for _ in range(2):
            x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            x = x ^ (x >> np.uint64(31))
    return ((x >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)


def basalt_material(begin: int, end: int, solid: bool = True) -> abi.Material:
    b = BASALT
    m = abi.Material()
    m.begin, m.end = begin, end
    m.eos = abi.EOS_TILLOTSON
    m.yielding = abi.YIELD_VON_MISES if solid else abi.YIELD_NONE
    m.fracture = abi.FRACTURE_SCALAR_GRADY_KIPP if solid else abi.FRACTURE_NONE
    for k in ("til_u0", "til_uiv", "til_ucv", "til_a", "til_b", "rho0", "til_A", "til_B", "til_alpha", "til_beta", "gamma",
              "shear_modulus", "elasticity_limit", "melt_energy"):
        setattr(m, k, b[k])
    mu, A = b["shear_modulus"], b["til_A"]
    m.young_modulus = mu * 9.0 * A / (3.0 * A + mu)  # Damage.cpp:44-48
    m.rho_min, m.rho_max = b["rho_range"]
    m.u_min, m.u_max = b["u_range"]
    m.d_min, m.d_max = b["d_range"]
    m.rho_small, m.u_small, m.d_small, m.s_small = b["rho_small"], b["u_small"], b["d_small"], b["s_small"]
    return m


def preset_config(solid: bool = True, adaptive_h: bool = True, correction_tensor: bool = True) -> abi.Config:
    """Collision preset (SphJob defaults minus SELF_GRAVITY) or the fluid-only variant of BASELINE configs[4]."""
    c = abi.Config()
    c.abi_version = abi.ABI_VERSION
    c.forces = abi.FORCE_PRESSURE | (abi.FORCE_SOLID_STRESS if solid else 0)
    c.flags = abi.FLAG_SUM_ONLY_UNDAMAGED | (abi.FLAG_ADAPTIVE_H if adaptive_h else 0) | (
        abi.FLAG_CORRECTION_TENSOR if (solid and correction_tensor) else 0)
    c.discretization, c.continuity_mode = 0, 0
    c.kernel_radius, c.av_alpha, c.av_beta = 2.0, 1.5, 3.0
    c.h_min, c.h_max = 1e-5, 1e10
    c.neigh_enforcing, c.neigh_lower, c.neigh_upper = 0.2, 25.0, 100.0
    c.criteria = abi.CRIT_COURANT | abi.CRIT_DIVERGENCE  # Presets.cpp:96-99
    c.courant, c.derivative_factor, c.divergence_factor = 0.2, 0.2, 0.005
    c.max_change = INF
    return c


def make_setup(n_particles: int, solid: bool = True, adaptive_h: bool = True, correction_tensor: bool = True) -> abi.RunSetup:
    grad, val = cubic_spline_lut()
    return abi.RunSetup(preset_config(solid, adaptive_h, correction_tensor), [basalt_material(0, n_particles, solid)], grad, val)


def basalt_sphere_state(n_target: int, radius: float = 5.0e4, solid: bool = True, seed: int = 1234, x_range=None,
                        total_hint: int = None, axis: int = 0) -> Dict[str, np.ndarray]:
    """Particle state of one basalt sphere on the hexagonal lattice (BASELINE configs[2..4]).

    Fields are seeded with smooth analytic functions of position (velocity ~ 50 m/s with compression and shear, 1 %
    density contrast, u ~ 1e4 J/kg, S ~ 1e7 Pa, damage 0) so AV, stress, EoS and damage terms all do real work.
    """
    pos3, h_lat = hexagonal_sphere(n_target, radius, x_range=x_range, axis=axis)
    n = len(pos3)
    b = BASALT
    x, y, z = (pos3[:, k] / radius for k in range(3))
    pos = np.empty((n, 4))
    pos[:, :3] = pos3
    pos[:, 3] = h_lat * b["eta"]
    vel = np.zeros((n, 4))
    vel[:, 0] = 50.0 * (-0.8 * x + 0.3 * y * z)
    vel[:, 1] = 50.0 * (0.5 * np.sin(3.0 * x) - 0.6 * y)
    vel[:, 2] = 50.0 * (0.4 * z * x - 0.7 * z + 0.2 * y)
    volume = 4.0 / 3.0 * math.pi * radius ** 3
    n_total = total_hint if total_hint is not None else n
    st: Dict[str, np.ndarray] = {
        "pos": pos, "vel": vel, "acc": np.zeros((n, 4)),
        "mass": np.full(n, b["rho0"] * volume / max(n_total, 1)),
        "rho": b["rho0"] * (1.0 + 0.01 * np.sin(4.0 * x) * np.cos(3.0 * y)),
        "u": 1.0e4 * (1.5 + np.cos(2.0 * z + x)),
        "p": np.zeros(n), "cs": np.zeros(n), "flag": np.zeros(n, np.uint32),
        "drho": np.zeros(n), "du": np.zeros(n),
    }
    if solid:
        S = np.empty((n, 5))
        S[:, 0] = 1.0e7 * np.sin(2.0 * x + y)
        S[:, 1] = 1.0e7 * np.cos(3.0 * y - z)
        S[:, 2] = 0.5e7 * np.sin(x * y * 4.0)
        S[:, 3] = 0.5e7 * np.cos(2.0 * z)
        S[:, 4] = 0.5e7 * np.sin(3.0 * x - 2.0 * z)
        st["S"], st["dS"] = S, np.zeros((n, 5))
        st["damage"], st["ddamage"], st["reduce"] = np.zeros(n), np.zeros(n), np.ones(n)
        # Weibull flaws, sampled variant (Damage.cpp:71-95); position-keyed hash instead of the reference's UniformRng
        A, mu = b["til_A"], b["shear_modulus"]
        cg = b["rayleigh"] * math.sqrt((A + 4.0 / 3.0 * mu) / b["rho0"])
        st["growth"] = cg / (2.0 * pos[:, 3])
        mw, kw = b["weibull_m"], b["weibull_k"]
        denom = 1.0 / (kw ** (1.0 / mw) * volume ** (1.0 / mw))
        size = float(max(n_total, 2))
        xr = _hash01(pos3, 0.05 * h_lat, seed, 1)
        p1 = -size * np.log1p(-xr)
        mult = math.exp(math.log(size)) - 1.0
        p2 = size * np.log1p(xr * mult)
        eps_min = denom * p1 ** (1.0 / mw)
        eps_max = denom * np.maximum(p1, p2) ** (1.0 / mw)
        from scipy.special import ndtri
        lam = math.log(size)  # Poisson(log N) flaws per particle, normal approximation keyed by position
        n_flaws = np.maximum(1, np.floor(lam + math.sqrt(lam) * ndtri(_hash01(pos3, 0.05 * h_lat, seed, 2)) + 0.5)).astype(np.uint32)
        eps_max = np.minimum(eps_max, n_flaws * eps_min)
        with np.errstate(divide="ignore", invalid="ignore"):
            m_zero = np.where(n_flaws == 1, 1.0, np.log(n_flaws) / np.log(np.maximum(eps_max / eps_min, 1.0 + 1e-12)))
        st["eps_min"], st["m_zero"], st["n_flaws"] = eps_min, m_zero, n_flaws
    return st
