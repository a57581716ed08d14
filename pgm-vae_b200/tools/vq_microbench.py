"""cfg4 microbench (BASELINE.json configs[3]): VQ assignment of N vectors (D=64) against a
K=8192 codebook, optionally followed by the EMA scatter.  Prints one JSON line.
    python pgm-vae_b200/tools/vq_microbench.py --n 16777216 --d 64 --k 8192 --prec tf32 --reps 5"""
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
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--k", type=int, default=8192)
    ap.add_argument("--prec", default="tf32", choices=["fp32", "tf32", "f16"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--clustered", action="store_true")
    ap.add_argument("--scatter", action="store_true", help="follow the assignment by the stand-alone EMA scatter")
    ap.add_argument("--fused", action="store_true", help="pgmvae_vq_assign_ema: scatter fused into the f16 kernel")
    a = ap.parse_args()
    ctx = _ffi.get_context(0)
    L = _ffi.lib()
    rng = np.random.default_rng(0)
    lim = np.sqrt(3.0 / a.d)
    e = rng.uniform(-lim, lim, (1, a.k, a.d)).astype(np.float32)
    if a.clustered:
        z = (e[0][rng.integers(0, a.k, a.n)] + 0.1 * rng.standard_normal((a.n, a.d))).astype(np.float32)[None]
    else:
        z = rng.standard_normal((1, a.n, a.d), dtype=np.float32)
    dz, de = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, e)
    idx = _ffi.DeviceArray(ctx, (1, a.n), np.int32)
    cnt, dw = _ffi.DeviceArray(ctx, (1, a.k)), _ffi.DeviceArray(ctx, (1, a.k, a.d))
    ctx.set_precision({"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32, "f16": _ffi.PREC_BF16}[a.prec])

    def run():
        if a.fused:
            _ffi.check(L.pgmvae_vq_assign_ema(ctx.h, None, dz.ptr, a.n * a.d, a.d, de.ptr, a.k * a.d, a.d, idx.ptr, a.n,
                                              cnt.ptr, a.k, dw.ptr, a.k * a.d, a.d, 1, a.n, a.d, a.k))
            return
        _ffi.check(L.pgmvae_vq_assign(ctx.h, None, dz.ptr, a.n * a.d, a.d, de.ptr, a.k * a.d, a.d, idx.ptr, a.n,
                                      None, None, 1, a.n, a.d, a.k))
        if a.scatter:
            _ffi.check(L.pgmvae_ema_stats(ctx.h, None, dz.ptr, a.n * a.d, a.d, idx.ptr, a.n, cnt.ptr, a.k, dw.ptr,
                                          a.k * a.d, a.d, 1, a.n, a.d, a.k))
    for _ in range(2):
        run()
    ctx.sync()
    ctx.timer_start()
    for _ in range(a.reps):
        run()
    ms = ctx.timer_stop_ms() / a.reps
    ctx.profile_begin()
    run()
    prof = ctx.profile_end()
    n = C.c_int(0)
    if a.prec != "fp32":
        _ffi.check(L.pgmvae_vq_assign_rescored(ctx.h, 1, a.k, C.byref(n)))
    flops = 2.0 * a.n * a.d * a.k
    print(json.dumps({"n": a.n, "d": a.d, "k": a.k, "prec": a.prec, "clustered": a.clustered, "scatter": a.scatter, "fused": a.fused, "sub": os.environ.get("PGMVAE_VQ_SUB", "2"),
                      "ms": ms, "vectors_per_s": a.n / (ms * 1e-3), "useful_tflops": flops / (ms * 1e-3) / 1e12,
                      "rescored_rows": n.value, "kernels": prof}))


if __name__ == "__main__":
    main()
