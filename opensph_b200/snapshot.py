"""Reader/writer of the "SPHSNAP1" particle snapshot container.

The format is a flat list of named arrays (f64 or u32, row-major [rows, ncomp]); it carries one particle state
in the reference's own particle order plus the run/material constants the engine needs.  Quantity names follow
the reference's QuantityId vocabulary: pos/vel/acc = POSITION value/dt/d2t as {x,y,z,h},
S = DEVIATORIC_STRESS {xx,yy,xy,xz,yz}, gradv/corr = SymmetricTensor {xx,yy,zz,xy,xz,yz}.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = b"SPHSNAP1"
_DTYPES = {0: np.float64, 1: np.uint32}
_CODES = {np.dtype(np.float64): 0, np.dtype(np.uint32): 1}


def read_snapshot(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not an SPHSNAP1 file")
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            name = f.read(32).split(b"\0", 1)[0].decode()
            dtype, ncomp, rows = struct.unpack("<IIQ", f.read(16))
            dt = np.dtype(_DTYPES[dtype])
            data = np.frombuffer(f.read(rows * ncomp * dt.itemsize), dtype=dt).copy()
            out[name] = data.reshape(rows, ncomp) if ncomp > 1 else data
    if "nbr_offsets" in out:  # u64 stored as pairs of u32 (little endian)
        off = out["nbr_offsets"].astype(np.uint64)
        out["nbr_offsets"] = off[:, 0] | (off[:, 1] << np.uint64(32))
    return out


def write_snapshot(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(arrays)))
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            if arr.dtype == np.uint64:  # keep the on-disk convention of the reference driver
                arr = np.stack([(arr & np.uint64(0xFFFFFFFF)).astype(np.uint32), (arr >> np.uint64(32)).astype(np.uint32)], axis=1)
            code = _CODES[arr.dtype]
            rows = arr.shape[0]
            ncomp = int(np.prod(arr.shape[1:])) if arr.ndim > 1 else 1
            f.write(name.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<IIQ", code, ncomp, rows))
            f.write(arr.tobytes())
