"""ctypes mirror of include/sphgpu.h (structs, enums) and helpers that build them from snapshot constants."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

ABI_VERSION = 1

# status codes
OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_OOM, E_STATE = 0, -1, -2, -3, -4, -5

FORCE_PRESSURE, FORCE_SOLID_STRESS = 1, 2
FLAG_CORRECTION_TENSOR, FLAG_SUM_ONLY_UNDAMAGED, FLAG_ADAPTIVE_H, FLAG_SOUND_SPEED_ENFORCING, FLAG_BALSARA = 1, 2, 4, 8, 16
FLAG_XSPH = 32
FLAG_DELTASPH = 64
FLAG_STRESS_AV = 128
EOS_NONE, EOS_IDEAL_GAS, EOS_TILLOTSON = 0, 1, 4
YIELD_NONE, YIELD_ELASTIC, YIELD_VON_MISES, YIELD_DUST = 0, 1, 2, 4
FRACTURE_NONE, FRACTURE_SCALAR_GRADY_KIPP = 0, 1
CRIT_COURANT, CRIT_DERIVATIVES, CRIT_ACCELERATION, CRIT_DIVERGENCE = 2, 4, 8, 16
LAYOUT_PACKED, LAYOUT_OPENSPH = 0, 1

# quantity ids: name -> (id, ncomp, dtype)
QUANTITIES: Dict[str, Tuple[int, int, type]] = {
    "POSITION": (0, 4, np.float64),
    "MASS": (1, 1, np.float64),
    "DENSITY": (2, 1, np.float64),
    "ENERGY": (3, 1, np.float64),
    "PRESSURE": (4, 1, np.float64),
    "SOUND_SPEED": (5, 1, np.float64),
    "DEVIATORIC_STRESS": (6, 5, np.float64),
    "DAMAGE": (7, 1, np.float64),
    "STRESS_REDUCING": (8, 1, np.float64),
    "VELOCITY_DIVERGENCE": (9, 1, np.float64),
    "VELOCITY_GRADIENT": (10, 6, np.float64),
    "CORRECTION_TENSOR": (11, 6, np.float64),
    "EPS_MIN": (12, 1, np.float64),
    "M_ZERO": (13, 1, np.float64),
    "EXPLICIT_GROWTH": (14, 1, np.float64),
    "N_FLAWS": (15, 1, np.uint32),
    "FLAG": (16, 1, np.uint32),
    "NEIGHBOR_CNT": (17, 1, np.uint32),
    "MATERIAL_ID": (18, 1, np.uint32),
    "VELOCITY_ROTATION": (19, 4, np.float64),
    "XSPH_VELOCITIES": (20, 4, np.float64),
    "DELTASPH_DENSITY_GRADIENT": (21, 4, np.float64),
    "AV_STRESS": (22, 6, np.float64),
    "INTERPARTICLE_SPACING_KERNEL": (23, 1, np.float64),
}

# snapshot array name -> (quantity, order)
SNAPSHOT_FIELDS: Dict[str, Tuple[str, int]] = {
    "pos": ("POSITION", 0), "vel": ("POSITION", 1), "acc": ("POSITION", 2),
    "mass": ("MASS", 0), "rho": ("DENSITY", 0), "drho": ("DENSITY", 1),
    "u": ("ENERGY", 0), "du": ("ENERGY", 1), "p": ("PRESSURE", 0), "cs": ("SOUND_SPEED", 0),
    "S": ("DEVIATORIC_STRESS", 0), "dS": ("DEVIATORIC_STRESS", 1),
    "damage": ("DAMAGE", 0), "ddamage": ("DAMAGE", 1), "reduce": ("STRESS_REDUCING", 0),
    "divv": ("VELOCITY_DIVERGENCE", 0), "gradv": ("VELOCITY_GRADIENT", 0), "corr": ("CORRECTION_TENSOR", 0),
    "eps_min": ("EPS_MIN", 0), "m_zero": ("M_ZERO", 0), "growth": ("EXPLICIT_GROWTH", 0),
    "n_flaws": ("N_FLAWS", 0), "flag": ("FLAG", 0), "ncnt": ("NEIGHBOR_CNT", 0), "xsph": ("XSPH_VELOCITIES", 0),
    "drho_grad": ("DELTASPH_DENSITY_GRADIENT", 0), "av_stress": ("AV_STRESS", 0), "wp": ("INTERPARTICLE_SPACING_KERNEL", 0),
}


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("forces", C.c_uint32), ("flags", C.c_uint32),
        ("discretization", C.c_uint32), ("continuity_mode", C.c_uint32), ("lut_entries", C.c_uint32),
        ("lut_grad", C.POINTER(C.c_double)), ("lut_value", C.POINTER(C.c_double)),
        ("kernel_radius", C.c_double), ("av_alpha", C.c_double), ("av_beta", C.c_double),
        ("h_min", C.c_double), ("h_max", C.c_double), ("neigh_enforcing", C.c_double),
        ("neigh_lower", C.c_double), ("neigh_upper", C.c_double),
        ("criteria", C.c_uint32), ("reserved0", C.c_uint32),
        ("courant", C.c_double), ("derivative_factor", C.c_double), ("divergence_factor", C.c_double),
        ("max_change", C.c_double),
    ]


class Material(C.Structure):
    _fields_ = [
        ("begin", C.c_uint32), ("end", C.c_uint32), ("eos", C.c_uint32), ("yielding", C.c_uint32),
        ("fracture", C.c_uint32), ("reserved0", C.c_uint32),
        ("til_u0", C.c_double), ("til_uiv", C.c_double), ("til_ucv", C.c_double), ("til_a", C.c_double),
        ("til_b", C.c_double), ("rho0", C.c_double), ("til_A", C.c_double), ("til_B", C.c_double),
        ("til_alpha", C.c_double), ("til_beta", C.c_double), ("gamma", C.c_double),
        ("shear_modulus", C.c_double), ("elasticity_limit", C.c_double), ("melt_energy", C.c_double),
        ("young_modulus", C.c_double),
        ("rho_min", C.c_double), ("rho_max", C.c_double), ("u_min", C.c_double), ("u_max", C.c_double),
        ("d_min", C.c_double), ("d_max", C.c_double),
        ("rho_small", C.c_double), ("u_small", C.c_double), ("d_small", C.c_double), ("s_small", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("neigh_min", C.c_uint32), ("neigh_max", C.c_uint32), ("neigh_mean", C.c_double),
        ("pair_count", C.c_uint64), ("gpu_ms", C.c_double), ("kernel_launches", C.c_uint32),
        ("reserved0", C.c_uint32),
    ]


class TimeStep(C.Structure):
    _fields_ = [("dt", C.c_double), ("criterion", C.c_uint32), ("reserved0", C.c_uint32)]


class Gravity(C.Structure):
    _fields_ = [
        ("opening_angle", C.c_double), ("multipole_order", C.c_int), ("leaf_size", C.c_uint32),
        ("constant", C.c_double), ("kernel_radius", C.c_double), ("lut_grad", C.POINTER(C.c_double)),
        ("lut_entries", C.c_uint32), ("reserved", C.c_uint32),
    ]


class GravityStats(C.Structure):
    _fields_ = [
        ("approximated", C.c_uint64), ("exact", C.c_uint64), ("nodes", C.c_uint32), ("groups", C.c_uint32),
        ("gpu_ms", C.c_double),
    ]


class Frozen(C.Structure):
    _fields_ = [
        ("flag_mask", C.c_uint64), ("has_domain", C.c_int), ("reserved", C.c_int), ("center", C.c_double * 3),
        ("radius", C.c_double), ("freeze_radius", C.c_double),
    ]


class Lattice(C.Structure):
    _fields_ = [
        ("center", C.c_double * 3), ("radius", C.c_double), ("particle_count", C.c_uint32), ("flags", C.c_uint32),
        ("eta", C.c_double), ("density", C.c_double), ("body_flag", C.c_uint32), ("reserved", C.c_uint32),
    ]


LATTICE_CENTER = 1
GRAVITY_CONSTANT = 6.67408e-11  # Constants::gravity (core/physics/Constants.h)


def gravity_table_cubic_spline(entries: int = 40000) -> np.ndarray:
    """Gradient table of GravityLutKernel for the cubic spline, as LutKernel<3>'s constructor fills it
    (core/sph/kernel/Kernel.h:85-101) from GravityKernel<CubicSpline<3>>::gradImpl (core/sph/kernel/GravityKernel.h:108-121):
    entries + 1 node values over q^2 in [0, 4]."""
    q_sqr = np.arange(entries + 1, dtype=np.float64) / (np.float64(entries) * 0.25)
    q = np.sqrt(q_sqr)
    out = np.empty(entries + 1)
    inner = q < 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        out[inner] = (4.0 / 3.0 * q[inner] - 6.0 / 5.0 * q[inner] ** 3 + 0.5 * q_sqr[inner] ** 2) / q[inner]
        o = ~inner
        out[o] = (8.0 / 3.0 * q[o] - 3.0 * q_sqr[o] + 6.0 / 5.0 * q[o] ** 3 - 1.0 / 6.0 * q_sqr[o] ** 2
                  - 1.0 / (15.0 * q_sqr[o])) / q[o]
    out[0] = 4.0 / 3.0
    return out


class RunSetup:
    """Config + materials + the arrays they point into (kept alive together)."""

    def __init__(self, cfg: Config, materials: Sequence[Material], lut_grad: np.ndarray, lut_value: np.ndarray):
        self.cfg = cfg
        self.materials = (Material * len(materials))(*materials)
        self.n_materials = len(materials)
        self.lut_grad = np.ascontiguousarray(lut_grad, dtype=np.float64)
        self.lut_value = np.ascontiguousarray(lut_value, dtype=np.float64)
        self.cfg.lut_grad = self.lut_grad.ctypes.data_as(C.POINTER(C.c_double))
        self.cfg.lut_value = self.lut_value.ctypes.data_as(C.POINTER(C.c_double))
        self.cfg.lut_entries = len(self.lut_grad) - 1
        self.xsph_eps = 1.0  # SPH_XSPH_EPSILON (used with FLAG_XSPH; Engine passes it to sphgpu_set_xsph_epsilon)
        self.deltasph_delta = 0.01  # SPH_DENSITY_DIFFUSION_DELTA, SPH_VELOCITY_DIFFUSION_ALPHA (used with FLAG_DELTASPH;
        self.deltasph_alpha = 0.01  # Engine passes them to sphgpu_set_deltasph)
        self.stress_av_exponent = 4.0  # SPH_AV_STRESS_EXPONENT, SPH_AV_STRESS_FACTOR (used with FLAG_STRESS_AV; Engine passes
        self.stress_av_factor = 0.04   # them to sphgpu_set_stress_av)

    @property
    def solid(self) -> bool:
        return bool(self.cfg.forces & FORCE_SOLID_STRESS)

    @property
    def has_damage(self) -> bool:
        return any(m.fracture != FRACTURE_NONE for m in self.materials)

    @property
    def has_reduce(self) -> bool:
        return any(m.yielding not in (YIELD_NONE,) for m in self.materials)


def _clip_inf(x: float) -> float:
    return float(x)


def setup_from_snapshot(snap: Dict[str, np.ndarray], lut: Dict[str, np.ndarray] = None) -> RunSetup:
    """Builds the engine configuration from the constants the reference driver stored (oracle/ref_driver.cpp:dumpState).

    `lut` supplies lut_grad / lut_val when the snapshot was written without the kernel tables."""
    if lut is None:
        lut = snap
    rp = snap["run_params"]
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.kernel_radius, cfg.av_alpha, cfg.av_beta = rp[0], rp[1], rp[2]
    cfg.forces = (FORCE_PRESSURE if rp[3] else 0) | (FORCE_SOLID_STRESS if rp[4] else 0)
    cfg.flags = ((FLAG_CORRECTION_TENSOR if (rp[5] and rp[4]) else 0) | (FLAG_SUM_ONLY_UNDAMAGED if rp[6] else 0)
                 | (FLAG_ADAPTIVE_H if rp[7] else 0) | (FLAG_SOUND_SPEED_ENFORCING if rp[8] else 0)
                 | (FLAG_BALSARA if (len(rp) > 25 and rp[25]) else 0) | (FLAG_XSPH if (len(rp) > 26 and rp[26]) else 0)
                 | (FLAG_DELTASPH if (len(rp) > 28 and rp[28]) else 0) | (FLAG_STRESS_AV if (len(rp) > 31 and rp[31]) else 0))
    cfg.continuity_mode = int(rp[9])
    cfg.discretization = int(rp[10])
    cfg.h_min, cfg.h_max = rp[11], rp[12]
    cfg.neigh_enforcing, cfg.neigh_lower, cfg.neigh_upper = rp[13], rp[14], rp[15]
    cfg.courant, cfg.derivative_factor, cfg.divergence_factor = rp[16], rp[17], rp[18]
    cfg.criteria = int(rp[19])
    cfg.max_change = rp[22]
    mats: List[Material] = []
    for (b, e), row in zip(snap["mat_range"].reshape(-1, 2), snap["mat_params"].reshape(-1, 32)):
        m = Material()
        m.begin, m.end = int(b), int(e)
        m.eos = int(row[0])
        (m.til_u0, m.til_uiv, m.til_ucv, m.til_a, m.til_b, m.rho0, m.til_A, m.til_B, m.til_alpha, m.til_beta,
         m.gamma) = row[1:12]
        m.yielding, m.fracture = int(row[12]), int(row[13])
        m.shear_modulus, m.elasticity_limit, m.melt_energy, m.young_modulus = row[14:18]
        m.rho_min, m.rho_max, m.u_min, m.u_max, m.d_min, m.d_max = row[18:24]
        m.rho_small, m.u_small, m.d_small, m.s_small = row[24:28]
        if m.yielding == YIELD_NONE:
            m.fracture = FRACTURE_NONE  # Factory::getMaterial: no rheology => EosMaterial (core/system/Factory.cpp:544-564)
        mats.append(m)
    setup = RunSetup(cfg, mats, lut["lut_grad"], lut["lut_val"])
    if len(rp) > 27:
        setup.xsph_eps = float(rp[27])
    if len(rp) > 30:
        setup.deltasph_delta, setup.deltasph_alpha = float(rp[29]), float(rp[30])
    if len(rp) > 33:
        setup.stress_av_exponent, setup.stress_av_factor = float(rp[32]), float(rp[33])
    return setup


def run_constants(snap: Dict[str, np.ndarray]) -> Dict[str, float]:
    rp = snap["run_params"]
    return {"max_dt": float(rp[20]), "initial_dt": float(rp[21]), "solver": int(rp[23]), "integrator": int(rp[24])}
