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


def iter_binary_csv(path: str, rows_per_chunk: int, nvar: Optional[int] = None):
    """Chunks y [<= rows_per_chunk, V] uint8 of a 0/1 CSV file without ever holding the file in memory (run.py:53: "design
    data pipeline for large dataset").  Fixed-width rows ("0,1,...,0\n": 2 V bytes) are cut straight out of the byte
    stream; anything else falls back to a line-wise parse of the chunk."""
    with open(path, "rb") as f:
        first = f.readline()
        if not first:
            return
        width = len(first)
        eol = 2 if first.endswith(b"\r\n") else 1
        v = (width - eol + 1) // 2
        if nvar is not None and v != nvar:
            raise ValueError(f"{path}: expected {nvar} columns, found {v}")
        f.seek(0)
        while True:
            raw = f.read(width * rows_per_chunk)
            if not raw:
                return
            if len(raw) % width:                     # last line without a newline
                raw += b"\n"
            buf = np.frombuffer(raw, dtype=np.uint8)
            ok = len(raw) % width == 0
            if ok:
                rows = buf.reshape(-1, width)
                digits = rows[:, 0:2 * v:2]
                ok = bool(np.all((digits == 48) | (digits == 49)) and np.all(rows[:, 1:2 * v - 1:2] == 44))
            if ok:
                yield np.ascontiguousarray(digits - 48)
            else:                                    # ragged lines: parse this chunk and whatever line it cut in two
                rest = f.readline()
                txt = (raw + rest).decode().strip().splitlines()
                yield (np.array([[float(t) for t in ln.split(",")] for ln in txt if ln], np.float32) != 0).astype(np.uint8)


def iter_array(y: np.ndarray, rows_per_chunk: int):
    for s in range(0, len(y), rows_per_chunk):
        yield y[s:s + rows_per_chunk]


class PinnedPrefetcher:
    """Runs a chunk iterator in a background thread and hands the chunks out in PINNED host buffers (a pool of
    `depth`), so that reading / parsing chunk i+1 overlaps the device work of chunk i and the H2D copies are DMA
    transfers.  ``for buf, rows in prefetcher: ...; prefetcher.release(buf)`` -- a buffer goes back to the pool once the
    device has consumed it."""

    def __init__(self, ctx, chunks, rows_per_chunk: int, nvar: int, depth: int = 3):
        import ctypes as C
        import queue
        import threading
        from pgmvae import _ffi
        self._ffi, self.ctx = _ffi, ctx
        self._ptrs, self._free, self._ready = [], queue.Queue(), queue.Queue(maxsize=depth)
        for _ in range(depth):
            hp = C.c_void_p()
            _ffi.check(_ffi.lib().pgmvae_malloc_host(ctx.h, rows_per_chunk * nvar, C.byref(hp)))
            self._ptrs.append(hp)
            self._free.put(np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(rows_per_chunk, nvar)))
        self._err = None

        def work():
            try:
                for ch in chunks:
                    ch = np.asarray(ch)
                    for s in range(0, len(ch), rows_per_chunk):
                        part = ch[s:s + rows_per_chunk]
                        buf = self._free.get()
                        buf[:len(part)] = part != 0 if part.dtype != np.uint8 else part
                        self._ready.put((buf, len(part)))
            except Exception as e:                   # surfaced on the consumer side
                self._err = e
            self._ready.put(None)
        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()

    def __iter__(self):
        while True:
            item = self._ready.get()
            if item is None:
                if self._err is not None:
                    raise self._err
                return
            yield item

    def release(self, buf):
        self._free.put(buf)

    def close(self):
        self._thread.join(timeout=30)
        for hp in self._ptrs:
            self._ffi.lib().pgmvae_free_host(self.ctx.h, hp)
        self._ptrs = []


def synthetic_binary(n: int, v: int, seed: int = 0) -> np.ndarray:
    """Column-wise Bernoulli data, densities U(0.02, 0.5) drawn once from default_rng(seed+1)."""
    dens = np.random.default_rng(seed + 1).uniform(0.02, 0.5, size=v)
    return (np.random.default_rng(seed).random((n, v)) < dens).astype(np.uint8)
