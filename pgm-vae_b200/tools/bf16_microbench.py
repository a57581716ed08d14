"""Kernel-only throughput of the bf16 tcgen05 GEMMs (csrc/dense_bf16.cu) at the cfg3 layer shapes, through the
operator ABI with the library's per-kernel CUDA-event profiler (the fp32 -> bf16 operand copies are separate,
named launches and are not counted).  Usage: python tools/bf16_microbench.py [G] [B]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pgmvae import _ffi  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    ctx = _ffi.get_context(0)
    ctx.set_precision(_ffi.PREC_BF16)
    L = _ffi.lib()
    rng = np.random.default_rng(0)
    res = []
    shapes = [(1556, 400), (400, 200), (200, 100), (100, 50), (50, 64), (64, 50), (400, 1556)]
    if len(sys.argv) > 3:
        shapes = [tuple(int(t) for t in a.split("x")) for a in sys.argv[3:]]
    for fin, fout in shapes:
        pin, pout = (fin + 7) // 8 * 8, (fout + 7) // 8 * 8
        x = _ffi.DeviceArray.from_numpy(ctx, rng.standard_normal((G, B, pin), dtype=np.float32))
        dy = _ffi.DeviceArray.from_numpy(ctx, rng.standard_normal((G, B, pout), dtype=np.float32))
        w = _ffi.DeviceArray.from_numpy(ctx, (rng.standard_normal((G, pin, pout), dtype=np.float32) * 0.05))
        b = _ffi.DeviceArray(ctx, (G, pout))
        out = _ffi.DeviceArray(ctx, (G, B, pout))
        dx = _ffi.DeviceArray(ctx, (G, B, pin))
        dw = _ffi.DeviceArray(ctx, (G, pin, pout))
        db = _ffi.DeviceArray(ctx, (G, pout))
        yv = _ffi.DeviceArray.from_numpy(ctx, (rng.random((B, pout)) < 0.2).astype(np.float32))
        acc = _ffi.DeviceArray(ctx, (4,), np.float64)

        def run():
            _ffi.check(L.pgmvae_dense_fwd(ctx.h, None, x.ptr, B * pin, pin, w.ptr, pin * pout, pout, b.ptr, pout, out.ptr,
                                          B * pout, pout, G, B, fin, fout, _ffi.ACT_SELU))
            if fout > 1000:
                _ffi.check(L.pgmvae_dense_fwd_sigmoid_mse(ctx.h, None, x.ptr, B * pin, pin, w.ptr, pin * pout, pout, b.ptr, pout,
                                                          yv.ptr, pout, out.ptr, B * pout, pout, None, acc.ptr,
                                                          G, 0, B, fin, fout, 1e-6))
            _ffi.check(L.pgmvae_dense_dgrad(ctx.h, None, dy.ptr, B * pout, pout, w.ptr, pin * pout, pout, x.ptr, B * pin, pin,
                                            None, None, 0, 0, 0.0, dx.ptr, B * pin, pin, G, B, fin, fout, _ffi.ACT_SELU))
            _ffi.check(L.pgmvae_dense_wgrad(ctx.h, None, x.ptr, B * pin, pin, dy.ptr, B * pout, pout, dw.ptr, pin * pout, pout,
                                            db.ptr, pout, G, B, fin, fout, -1))
        for orient in (("d", "t") if max(fin, fout) > 1000 else ("",)):
            if orient:
                os.environ["PGMVAE_WGRAD_ORIENT"] = orient
            else:
                os.environ.pop("PGMVAE_WGRAD_ORIENT", None)
            run()
            ctx.sync()
            ctx.profile_begin()
            for _ in range(3):
                run()
            for k in ctx.profile_end():
                if "bf16" not in k["name"] or k["name"] in ("f32_to_bf16", "bf16_shadow"):
                    continue
                ms = k["ms"] / k["launches"]
                res.append({"shape": f"{fin}->{fout}", "G": G, "B": B, "orient": orient, "kernel": k["name"], "ms": ms,
                            "TFLOPs": k["flops"] / k["launches"] / (ms * 1e-3) / 1e12 if k["flops"] else None,
                            "GBps": k["bytes"] / k["launches"] / (ms * 1e-3) / 1e9})
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
