"""ctypes wrapper of the plain-C oracle (oracle/liboracle_port.so). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from opensph_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liboracle_port.so")

_D = C.POINTER(C.c_double)
_U = C.POINTER(C.c_uint32)


class OrcState(C.Structure):
    _fields_ = [("n", C.c_uint32), ("pad0", C.c_uint32)] + [(k, _D) for k in (
        "pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "reduce", "damage", "ddamage",
        "eps_min", "m_zero", "growth")] + [(k, _U) for k in ("n_flaws", "flag", "ncnt")] + [(k, _D) for k in (
        "divv", "gradv", "corr", "acc_pred", "drho_pred", "du_pred", "dS_pred", "ddamage_pred", "xsph")] + [("xsph_eps", C.c_double), ("drho_grad", _D),
                                                                                     ("deltasph_delta", C.c_double), ("deltasph_alpha", C.c_double),
                                                                                     ("av_stress", _D), ("wp", _D),
                                                                                     ("stress_av_exponent", C.c_double), ("stress_av_factor", C.c_double)]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_find_neighbours.restype = C.c_uint64
        _lib.orc_timestep.restype = C.c_double
    return _lib


_F64 = ("pos", "vel", "acc", "mass", "rho", "drho", "u", "du", "p", "cs", "S", "dS", "reduce", "damage", "ddamage",
        "eps_min", "m_zero", "growth", "divv", "gradv", "corr", "xsph", "drho_grad", "av_stress", "wp")
_U32 = ("n_flaws", "flag", "ncnt")
_PRED = {"acc_pred": "acc", "drho_pred": "drho", "du_pred": "du", "dS_pred": "dS", "ddamage_pred": "ddamage"}


class OraclePort:
    """Holds one particle state (copied from a snapshot dict) and runs the C restatement on it."""

    def __init__(self, snap: Dict[str, np.ndarray], setup: Optional[abi.RunSetup] = None):
        self.setup = setup or abi.setup_from_snapshot(snap)
        self.a: Dict[str, np.ndarray] = {}
        n = len(snap["mass"])
        for k in _F64:
            if k in snap:
                self.a[k] = np.ascontiguousarray(snap[k], dtype=np.float64).copy()
        for k in _U32:
            if k in snap:
                self.a[k] = np.ascontiguousarray(snap[k], dtype=np.uint32).copy()
        if "ncnt" not in self.a:
            self.a["ncnt"] = np.zeros(n, np.uint32)
        if "flag" not in self.a:
            self.a["flag"] = np.zeros(n, np.uint32)
        if "divv" not in self.a:
            self.a["divv"] = np.zeros(n)
        if self.setup.solid:
            self.a.setdefault("gradv", np.zeros((n, 6)))
            self.a.setdefault("corr", np.tile(np.array([1., 1., 1., 0., 0., 0.]), (n, 1)))
        for pk, k in _PRED.items():
            if k in self.a:
                self.a[pk] = np.zeros_like(self.a[k])
        self.n = n
        self.state = OrcState()
        self.state.n = n
        if self.setup.cfg.flags & abi.FLAG_XSPH:
            self.a.setdefault("xsph", np.zeros((n, 4)))
        self.state.xsph_eps = self.setup.xsph_eps
        if self.setup.cfg.flags & abi.FLAG_DELTASPH:
            self.a.setdefault("drho_grad", np.zeros((n, 4)))
        self.state.deltasph_delta = self.setup.deltasph_delta
        self.state.deltasph_alpha = self.setup.deltasph_alpha
        if self.setup.cfg.flags & abi.FLAG_STRESS_AV:
            self.a.setdefault("av_stress", np.zeros((n, 6)))
        self.state.stress_av_exponent = self.setup.stress_av_exponent
        self.state.stress_av_factor = self.setup.stress_av_factor
        for name, _ in OrcState._fields_[2:]:
            if name in ("xsph_eps", "deltasph_delta", "deltasph_alpha", "stress_av_exponent", "stress_av_factor"):
                continue
            arr = self.a.get(name)
            if arr is None:
                setattr(self.state, name, None)
            elif arr.dtype == np.uint32:
                setattr(self.state, name, arr.ctypes.data_as(_U))
            else:
                setattr(self.state, name, arr.ctypes.data_as(_D))
        self.last_dt = C.c_double(0.0)

    def _args(self):
        return C.byref(self.state), C.byref(self.setup.cfg), self.setup.materials, C.c_uint32(self.setup.n_materials)

    def integrate(self) -> None:
        lib().orc_integrate(*self._args())

    def frozen(self, flags=(), domain=None) -> None:
        """FrozenParticles::finalize after integrate(): bodies `flags`, domain = (centre, radius, freeze_radius)."""
        mask = 0
        for b in flags:
            mask |= 1 << int(b)
        c = np.zeros(3) if domain is None else np.ascontiguousarray(domain[0], dtype=np.float64)
        lib().orc_frozen(C.byref(self.state), C.c_int(1 if self.setup.solid else 0), C.c_uint64(mask), C.c_int(0 if domain is None else 1),
                         c.ctypes.data_as(_D), C.c_double(0.0 if domain is None else domain[1]), C.c_double(0.0 if domain is None else domain[2]))

    def predict(self, dt: float) -> None:
        lib().orc_predict(*self._args(), C.c_double(dt))

    def correct(self, dt: float) -> None:
        lib().orc_correct(*self._args(), C.c_double(dt))

    def euler(self, dt: float) -> None:
        lib().orc_euler(*self._args(), C.c_double(dt))

    def timestep(self, max_dt: float):
        crit = C.c_uint32(0)
        dt = lib().orc_timestep(*self._args(), C.c_double(max_dt), C.byref(self.last_dt), C.byref(crit))
        return float(dt), int(crit.value)

    def step_pc(self, dt: float, max_dt: float):
        """ITimeStepping::step with PredictorCorrector (core/timestepping/TimeStepping.cpp:34-75,324-346)."""
        self.predict(dt)
        self.integrate()
        self.correct(dt)
        return self.timestep(max_dt)

    def step_euler(self, dt: float, max_dt: float):
        self.integrate()
        self.euler(dt)
        return self.timestep(max_dt)

    def neighbours(self):
        off = np.zeros(self.n + 1, np.uint64)
        total = lib().orc_find_neighbours(C.byref(self.state), C.byref(self.setup.cfg),
                                          off.ctypes.data_as(C.POINTER(C.c_uint64)), None, C.c_uint64(0))
        idx = np.zeros(int(total), np.uint32)
        lib().orc_find_neighbours(C.byref(self.state), C.byref(self.setup.cfg),
                                  off.ctypes.data_as(C.POINTER(C.c_uint64)), idx.ctypes.data_as(_U), C.c_uint64(int(total)))
        return off, idx


def build_lut(entries: int = 40000, radius: float = 2.0):
    grad = np.zeros(entries + 1)
    val = np.zeros(entries + 1)
    lib().orc_build_lut(grad.ctypes.data_as(_D), val.ctypes.data_as(_D), C.c_uint32(entries), C.c_double(radius))
    return grad, val


# ---- self-gravity (oracle/sph_oracle.c: orc_gravity_*) -------------------------------------------------------------
def gravity_lut(entries: int = 40000) -> np.ndarray:
    out = np.empty(entries + 1)
    lib().orc_build_gravity_lut(out.ctypes.data_as(_D), C.c_uint32(entries))
    return out


def gravity_brute(pos: np.ndarray, mass: np.ndarray, G: float, lut_grad: Optional[np.ndarray], radius: float) -> np.ndarray:
    """BruteForceGravity restated in C: accelerations [n, 3]."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    n = len(mass)
    acc = np.zeros((n, 3))
    if lut_grad is None:
        lut_grad, radius = np.zeros(2), 0.0
    lut_grad = np.ascontiguousarray(lut_grad, dtype=np.float64)
    lib().orc_gravity_brute(C.c_uint32(n), pos.ctypes.data_as(_D), mass.ctypes.data_as(_D), C.c_double(G),
                            lut_grad.ctypes.data_as(_D), C.c_uint32(len(lut_grad) - 1), C.c_double(radius),
                            acc.ctypes.data_as(_D))
    return acc


def gravity_moments(pos: np.ndarray, mass: np.ndarray):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    com, mom = np.zeros(3), np.zeros(13)
    lib().orc_gravity_moments(C.c_uint32(len(mass)), pos.ctypes.data_as(_D), mass.ctypes.data_as(_D), com.ctypes.data_as(_D),
                              mom.ctypes.data_as(_D))
    return com, mom


def gravity_multipole(com: np.ndarray, mom: np.ndarray, order: int, point: np.ndarray) -> np.ndarray:
    com = np.ascontiguousarray(com, dtype=np.float64)
    mom = np.ascontiguousarray(mom, dtype=np.float64)
    point = np.ascontiguousarray(point[:3], dtype=np.float64)
    acc = np.zeros(3)
    lib().orc_gravity_multipole(com.ctypes.data_as(_D), mom.ctypes.data_as(_D), C.c_int(order), point.ctypes.data_as(_D),
                                acc.ctypes.data_as(_D))
    return acc


# ---- initial conditions (oracle/sph_oracle.c: orc_hexagonal_sphere) --------------------------------------------------
def hexagonal_sphere(n: int, centre, radius: float, centred: bool = True, eta: float = 1.3, rho0: float = 2700.0):
    """InitialConditions::addMonolithicBody (SphericalDomain, HexagonalPacking) restated in C: (pos [m, 4], mass [m])."""
    lib().orc_hexagonal_sphere.restype = C.c_uint32
    c = np.ascontiguousarray(centre, dtype=np.float64)
    args = (C.c_uint32(n), c.ctypes.data_as(_D), C.c_double(radius), C.c_int(1 if centred else 0), C.c_double(eta), C.c_double(rho0))
    m = lib().orc_hexagonal_sphere(*args, None, None, C.c_uint32(0))
    pos, mass = np.zeros((m, 4)), np.zeros(m)
    lib().orc_hexagonal_sphere(*args, pos.ctypes.data_as(_D), mass.ctypes.data_as(_D), C.c_uint32(m))
    return pos, mass
