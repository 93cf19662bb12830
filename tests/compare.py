"""Parity metric shared by all tests: |a-b| <= tol * max(|a|,|b|, s_q), s_q = floor * max|q|."""
from __future__ import annotations

import numpy as np


def rel_err(a: np.ndarray, b: np.ndarray, floor: float = 1e-6) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = floor * max(np.abs(a).max(), np.abs(b).max(), 1e-300)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    return float((np.abs(a - b) / denom).max())


def assert_close(name: str, a, b, tol: float = 1e-10, floor: float = 1e-6) -> None:
    e = rel_err(a, b, floor)
    assert e <= tol, f"{name}: relative error {e:.3e} > {tol:.1e}"
