"""HBM-bound stages on their own: the stand-alone EMA scatter + update at the cfg4
shape and the PLL histogram at the cfg5 shape (1556 variables).  Prints one JSON line.
    python pgm-vae_b200/tools/hbm_microbench.py [--n 16777216] [--b 65536]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgmvae import _ffi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 24, help="vectors for the EMA scatter (D=64, K=8192)")
    ap.add_argument("--b", type=int, default=65536, help="samples for the PLL histogram (V=1556, K=512)")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    ctx = _ffi.get_context(0)
    L = _ffi.lib()
    rng = np.random.default_rng(0)
    out = {}
    # ---- EMA statistics scatter + update, cfg4: N x 64 vectors, 8192 codes
    D, K, N = 64, 8192, a.n
    z = _ffi.DeviceArray.from_numpy(ctx, rng.standard_normal((1, N, D), dtype=np.float32))
    idx = _ffi.DeviceArray.from_numpy(ctx, rng.integers(0, K, (1, N)).astype(np.int32))
    cnt, dw = _ffi.DeviceArray(ctx, (1, K)), _ffi.DeviceArray(ctx, (1, K, D))
    bc, bw = _ffi.DeviceArray(ctx, (1, K)), _ffi.DeviceArray(ctx, (1, K, D))
    ec, ew, e = _ffi.DeviceArray(ctx, (1, K)), _ffi.DeviceArray(ctx, (1, K, D)), _ffi.DeviceArray(ctx, (1, K, D))

    def ema():
        _ffi.check(L.pgmvae_ema_stats(ctx.h, None, z.ptr, N * D, D, idx.ptr, N, cnt.ptr, K, dw.ptr, K * D, D, 1, N, D, K))
        _ffi.check(L.pgmvae_ema_apply(ctx.h, None, cnt.ptr, dw.ptr, bc.ptr, bw.ptr, ec.ptr, ew.ptr, e.ptr, 1, K, D, D,
                                      0.99, 1e-5, 1, 1))
    ema(); ctx.sync()
    ctx.profile_begin()
    for _ in range(a.reps):
        ema()
    out["ema"] = {"n": N, "d": D, "k": K, "kernels": ctx.profile_end()}
    del z, idx
    # ---- PLL histogram, cfg5 shape: V=1556 variables, K=512 codes, B samples per call
    V, K2, B = 1556, 512, a.b
    idx2 = _ffi.DeviceArray.from_numpy(ctx, rng.integers(0, K2, (V, B)).astype(np.int32))
    y = _ffi.DeviceArray.from_numpy(ctx, (rng.random((B, V)) < 0.2).astype(np.uint8))
    n1 = _ffi.DeviceArray(ctx, (V, K2), np.uint64)
    n0 = _ffi.DeviceArray(ctx, (V, K2), np.uint64)

    def pll():
        _ffi.check(L.pgmvae_pll_count(ctx.h, None, idx2.ptr, B, y.ptr, V, 0, n1.ptr, n0.ptr, V, B, K2))
    pll(); ctx.sync()
    ctx.profile_begin()
    for _ in range(a.reps):
        pll()
    out["pll_count"] = {"v": V, "k": K2, "b": B, "kernels": ctx.profile_end()}
    assert int(n1.numpy().sum() + n0.numpy().sum()) == (a.reps + 1) * V * B
    for sec in out.values():
        for k in sec["kernels"]:
            k["GBps"] = k["bytes"] / max(k["ms"], 1e-9) / 1e6
            k["ms_per_launch"] = k["ms"] / k["launches"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
