"""Data ingest for the hot path: 0/1 CSV files (reference data/trw/*.data, run.py:52-56)
-> uint8 matrix y [N,V].  The reference parses rows through tf.data and then materialises
x [N,V,V-1]; here only y is produced (the kernels read nothing else).

Also holds the synthetic generator of the benchmark shapes."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PACKED_DIR = os.path.join(os.path.dirname(_HERE), "data", "packed")


def parse_binary_csv(path: str, nvar: Optional[int] = None) -> np.ndarray:
    """Fast path for fixed-width rows "0,1,0,...": one byte in two is a digit."""
    raw = open(path, "rb").read().replace(b"\r", b"")
    if not raw:
        return np.zeros((0, nvar or 0), np.uint8)
    if not raw.endswith(b"\n"):
        raw += b"\n"
    first = raw.index(b"\n")
    width = first + 1
    v = (first + 1) // 2
    buf = np.frombuffer(raw, dtype=np.uint8)
    if len(raw) % width == 0:
        rows = buf.reshape(-1, width)
        digits, seps = rows[:, 0:first:2], rows[:, 1:first:2]
        if (np.all((digits == 48) | (digits == 49)) and np.all(seps == 44) and np.all(rows[:, first] == 10)
                and (nvar is None or v == nvar)):
            return np.ascontiguousarray(digits - 48)
    y = np.loadtxt(path, delimiter=",", dtype=np.float32, ndmin=2)      # general fallback
    if nvar is not None and y.shape[1] != nvar:
        raise ValueError(f"{path}: expected {nvar} columns, found {y.shape[1]}")
    return (y != 0).astype(np.uint8)


def load_split(name: str, tvt: str, nvar: Optional[int] = None, root: Optional[str] = None) -> np.ndarray:
    """y [N,V] uint8 of data/trw/{name}.{tvt}.data (run.py:54), searched in `root`,
    $PGMVAE_DATA, ./data/trw, then the bit-packed copy shipped for the smallest dataset."""
    for base in [root, os.environ.get("PGMVAE_DATA"), os.path.join(os.curdir, "data", "trw"),
                 os.path.join(os.path.dirname(_HERE), "data", "trw")]:
        if base:
            p = os.path.join(base, f"{name}.{tvt}.data")
            if os.path.exists(p):
                return parse_binary_csv(p, nvar)
    packed = os.path.join(PACKED_DIR, f"{name}.npz")
    if os.path.exists(packed):
        with np.load(packed) as z:
            v = int(z["vars"])
            n = int(z[f"{tvt}_rows"])
            return np.unpackbits(z[tvt], axis=None)[: n * v].reshape(n, v).astype(np.uint8)
    raise FileNotFoundError(f"data/trw/{name}.{tvt}.data not found (set PGMVAE_DATA to the data/trw directory)")


def synthetic_binary(n: int, v: int, seed: int = 0) -> np.ndarray:
    """Column-wise Bernoulli data, densities U(0.02, 0.5) drawn once from default_rng(seed+1)."""
    dens = np.random.default_rng(seed + 1).uniform(0.02, 0.5, size=v)
    return (np.random.default_rng(seed).random((n, v)) < dens).astype(np.uint8)
