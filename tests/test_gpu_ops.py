"""GPU parity tests, operator level: every kernel called through the C-ABI (via the host
mirror core.dense / core.quantizer and pgmvae._ffi) against the CPU oracle on the same
seeded inputs.  Tolerances: indices bit-exact where the top-2 distance gap exceeds 1e-5
(lowest index on ties); floating point within 1e-3 relative (north_star), tightened to
what fp32 summation-order differences allow."""
import ctypes as C

import os

import numpy as np
import pytest
import torch

import pgmvae_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("V,B,fin,fout,act", [
    (16, 256, 15, 15, "selu"), (16, 53, 12, 4, "selu"), (3, 1, 4, 12, "selu"), (5, 200, 68, 50, "selu"),
    (2, 129, 130, 67, "sigmoid"), (4, 64, 16, 24, None), (69, 300, 20, 16, "selu"),
])
def test_fatdense_forward(ctx, V, B, fin, fout, act):
    from core.dense import FatDense
    rng = np.random.default_rng(V * 1000 + B)
    x = rng.standard_normal((V, B, fin)).astype(np.float32)
    layer = FatDense(fout, activation=act, kernel_initializer="he_uniform")
    layer.build(x.shape)
    layer.bias = rng.standard_normal((V, 1, fout)).astype(np.float32) * 0.1
    got = layer(x).numpy()
    exp = O.fatdense_call(torch.from_numpy(x), torch.from_numpy(layer.kernel), torch.from_numpy(layer.bias), act).numpy()
    assert got.shape == (V, B, fout)
    np.testing.assert_allclose(got, exp, rtol=2e-5, atol=2e-6)


def test_fatdense_fts_subset(ctx):
    from core.dense import FatDense
    rng = np.random.default_rng(7)
    V, B, fin, fout = 9, 40, 8, 5
    layer = FatDense(fout, activation="selu", kernel_initializer="glorot_uniform")
    layer.build((V, B, fin))
    fts = np.array([7, 2, 2, 0])
    x = rng.standard_normal((len(fts), B, fin)).astype(np.float32)
    got = layer(x, fts=fts).numpy()
    exp = O.fatdense_call(torch.from_numpy(x), torch.from_numpy(layer.kernel), torch.from_numpy(layer.bias), "selu",
                          fts=torch.from_numpy(fts)).numpy()
    np.testing.assert_allclose(got, exp, rtol=2e-5, atol=2e-6)


def test_fatdense_accepts_torch_cuda_carrier(ctx):
    """PyTorch is only a tensor carrier: a CUDA tensor crosses the ABI by pointer."""
    from core.dense import FatDense
    if not torch.cuda.is_available():
        pytest.skip("torch CUDA not available")
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 10, 6)).astype(np.float32)
    layer = FatDense(4, activation="selu")
    layer.build(x.shape)
    xt = torch.from_numpy(x).cuda()
    torch.cuda.synchronize()
    got = layer(xt).numpy()
    np.testing.assert_allclose(got, layer(x).numpy(), rtol=0, atol=0)


def _assign_case(V, B, D, K, seed, clustered=False):
    rng = np.random.default_rng(seed)
    lim = np.sqrt(3.0 / (V * D))
    emb = rng.uniform(-lim, lim, (V, D, K)).astype(np.float32)
    z = rng.standard_normal((V, B, D)).astype(np.float32) * (0.2 if not clustered else 1.0)
    if clustered:
        pick = rng.integers(0, K, (V, B))
        z = (np.take_along_axis(emb.transpose(0, 2, 1), pick[..., None], 1) + 0.1 * lim * rng.standard_normal((V, B, D))
             ).astype(np.float32)
    return z, emb


@pytest.mark.parametrize("V,B,D,K,clustered", [
    (16, 256, 4, 32, False), (16, 53, 4, 32, True), (69, 512, 16, 128, False), (3, 1000, 64, 512, True),
    (1, 4096, 64, 1024, False), (2, 130, 10, 7, False), (5, 77, 30, 1, False),
])
@pytest.mark.parametrize("prec", ["fp32", "tf32", "f16"])
def test_vq_assign_indices(ctx, V, B, D, K, clustered, prec):
    from core.quantizer import VectorQuantizer
    from pgmvae import _ffi
    z, emb = _assign_case(V, B, D, K, seed=V + B + D + K, clustered=clustered)
    layer = VectorQuantizer(D, K, 0.25, V)
    layer.embeddings = emb
    ctx.set_precision({"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32, "f16": _ffi.PREC_BF16}[prec])
    try:
        onehot = layer(z, code_only=True)
    finally:
        ctx.set_precision(_ffi.PREC_FP32)
    got = layer.last_indices.numpy()
    idx, gap = O.vq_assign(z, emb)
    idx, gap = idx.numpy(), gap.numpy()
    safe = gap > 1e-5
    assert safe.mean() > 0.9 or K == 1
    np.testing.assert_array_equal(got[safe], idx[safe])
    assert onehot.shape == (V, B, K) and np.array_equal(onehot.argmax(-1), got)
    # rows inside the 1e-5 band must still pick a near-minimal code
    d = O.vq_distances(torch.from_numpy(z), torch.from_numpy(emb)).numpy()
    picked = np.take_along_axis(d, got[..., None].astype(np.int64), 2)[..., 0]
    assert np.all(picked - d.min(2) <= 1e-4)


def test_vq_assign_ties_pick_lowest_index(ctx):
    from core.quantizer import VectorQuantizer
    V, B, D, K = 2, 33, 8, 12
    rng = np.random.default_rng(0)
    emb = rng.uniform(-0.5, 0.5, (V, D, K)).astype(np.float32)
    emb[:, :, 9] = emb[:, :, 3]
    emb[:, :, 5] = emb[:, :, 3]                     # codes 3, 5, 9 identical
    z = np.repeat(emb[:, :, 3][:, None, :], B, 1).astype(np.float32)
    layer = VectorQuantizer(D, K, 0.25, V)
    layer.embeddings = emb
    layer(z, code_only=True)
    assert np.all(layer.last_indices.numpy() == 3)


@pytest.mark.parametrize("ema", [True, False])
def test_vq_layer_outputs_and_losses(ctx, ema):
    from core.quantizer import VectorQuantizer, VectorQuantizerEMA
    V, B, D, K = 6, 97, 5, 11
    z, emb = _assign_case(V, B, D, K, seed=11)
    layer = VectorQuantizerEMA(D, K, 0.25, 0.99, V) if ema else VectorQuantizer(D, K, 0.25, V)
    layer.embeddings = emb
    out = layer(z, training=False).numpy()
    om = O.OracleVqVAE([3, 3, 3, 3], V, D, K, cost=0.25, ema=ema,
                       params={**O.init_params([3, 3, 3, 3], V, D, K), "vq.embeddings": torch.from_numpy(emb)})
    exp = om.vq_layer(torch.from_numpy(z), training=False).numpy()
    np.testing.assert_allclose(out, exp, rtol=1e-6, atol=1e-7)
    ref_loss = float(om.losses[0].detach())
    assert abs(layer.losses[0] - ref_loss) <= 1e-5 * abs(ref_loss)


def test_ema_update_three_steps(ctx):
    """counts / dw scatter-add and the debiased EMA + Laplace normalisation, three steps."""
    from core.quantizer import VectorQuantizerEMA
    V, B, D, K = 4, 200, 6, 9
    z0, emb = _assign_case(V, B, D, K, seed=21)
    layer = VectorQuantizerEMA(D, K, 0.25, 0.9, V)
    layer.build((V, B, D))
    layer.embeddings = emb
    layer.ema_w = emb
    om = O.OracleVqVAE([3, 3, 3, 3], V, D, K, cost=0.25, decay=0.9, ema=True,
                       params={**O.init_params([3, 3, 3, 3], V, D, K), "vq.embeddings": torch.from_numpy(emb)})
    rng = np.random.default_rng(5)
    for step in range(3):
        z = (z0 + 0.05 * rng.standard_normal(z0.shape)).astype(np.float32)
        layer(z, training=True)
        om.vq_layer(torch.from_numpy(z), training=True)
        st = om.ema_state
        np.testing.assert_allclose(layer.ema_cluster_size, st.ema_cluster_size.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(layer.ema_w, st.ema_w.numpy(), rtol=1e-4, atol=1e-6)
        assert rel_err(layer.embeddings, om.p["vq.embeddings"].numpy()) < RTOL


@pytest.mark.parametrize("G,B,D,K", [(3, 5000, 16, 128), (1, 20000, 64, 1024), (2, 300, 4, 3)])
def test_ema_stats_scatter(ctx, G, B, D, K):
    from pgmvae import _ffi
    rng = np.random.default_rng(G + B)
    z = rng.standard_normal((G, B, D)).astype(np.float32)
    idx = rng.integers(0, K, (G, B)).astype(np.int32)
    idx[:, : B // 3] = idx[:, :1]                              # skewed code usage
    dz, di = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, idx)
    cnt, dw = _ffi.DeviceArray(ctx, (G, K)), _ffi.DeviceArray(ctx, (G, K, D))
    _ffi.check(_ffi.lib().pgmvae_ema_stats(ctx.h, None, dz.ptr, B * D, D, di.ptr, B, cnt.ptr, K, dw.ptr, K * D, D,
                                           G, B, D, K))
    c, w = O.ema_stats(z, idx, K)
    np.testing.assert_array_equal(cnt.numpy(), c.numpy())
    np.testing.assert_allclose(dw.numpy(), w.numpy().transpose(0, 2, 1), rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("V,B,K", [(16, 1000, 32), (69, 5000, 128), (3, 7, 512), (40, 2049, 8)])
def test_pll_count_cpt_reduce(ctx, V, B, K):
    from pgmvae import _ffi
    rng = np.random.default_rng(V + B + K)
    idx = rng.integers(0, K, (V, B)).astype(np.int32)
    y = O.synthetic_binary(B, V, seed=V)
    di, dy = _ffi.DeviceArray.from_numpy(ctx, idx), _ffi.DeviceArray.from_numpy(ctx, y)
    n1, n0 = _ffi.DeviceArray(ctx, (V, K), np.uint64), _ffi.DeviceArray(ctx, (V, K), np.uint64)
    L = _ffi.lib()
    _ffi.check(L.pgmvae_pll_count(ctx.h, None, di.ptr, B, dy.ptr, V, 0, n1.ptr, n0.ptr, V, B, K))
    h1, h0 = np.zeros((V, K), np.uint64), np.zeros((V, K), np.uint64)
    for v in range(V):
        np.add.at(h1[v], idx[v][y[:, v] != 0], 1)
        np.add.at(h0[v], idx[v][y[:, v] == 0], 1)
    np.testing.assert_array_equal(n1.numpy(), h1)
    np.testing.assert_array_equal(n0.numpy(), h0)
    dist = _ffi.DeviceArray(ctx, (V, K), np.float64)
    _ffi.check(L.pgmvae_cpt(ctx.h, None, n1.ptr, n0.ptr, dist.ptr, V * K))
    exp_dist = (h1.astype(np.float64) + 0.8) / (h1.astype(np.float64) + h0.astype(np.float64) + 1.6)
    np.testing.assert_allclose(dist.numpy(), exp_dist, rtol=1e-15)
    out = _ffi.DeviceArray(ctx, (1,), np.float64)
    _ffi.check(L.pgmvae_pll_reduce(ctx.h, None, n1.ptr, n0.ptr, dist.ptr, V * K, out.ptr))
    exp = O.pll_from_counts(h1, h0, exp_dist, B)
    assert abs(out.numpy()[0] / B - exp) <= 1e-12 * abs(exp)


def test_adam_step(ctx):
    from pgmvae import _ffi
    rng = np.random.default_rng(3)
    n = 10007
    p, g = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32) * 1e-3
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    dp, dg, dm, dv = (_ffi.DeviceArray.from_numpy(ctx, a) for a in (p, g, m, v))
    pe = p.copy()
    for t in range(1, 4):
        alpha = np.float32(1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t))
        _ffi.check(_ffi.lib().pgmvae_adam_step(ctx.h, None, dp.ptr, dg.ptr, dm.ptr, dv.ptr, n, alpha, 0.9, 0.999, 1e-7))
        m += (g - m) * np.float32(0.1)
        v += (g * g - v) * np.float32(0.001)
        pe -= (m * alpha) / (np.sqrt(v) + np.float32(1e-7))
    np.testing.assert_allclose(dp.numpy(), pe, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dm.numpy(), m, rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("G,B,D,K", [(1, 8192, 64, 8192), (3, 1000, 64, 512), (69, 4096, 16, 128), (2, 300, 128, 600),
                                     (16, 256, 4, 32), (2, 513, 40, 100)])
def test_vq_assign_tensor_core_equals_fp32_path(ctx, G, B, D, K):
    """tcgen05 (tf32) assignment + fp32 re-scoring of the ambiguous rows must reproduce the exact-fp32
    CUDA-core kernel index for index (and therefore the oracle outside the 1e-5 band)."""
    from pgmvae import _ffi
    import ctypes as C
    rng = np.random.default_rng(G + B + D + K)
    z = rng.standard_normal((G, B, D)).astype(np.float32)
    e = rng.uniform(-1, 1, (G, K, D)).astype(np.float32) * np.float32(np.sqrt(3.0 / D))
    dz, de = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, e)
    L = _ffi.lib()
    out = {}
    nres = {}
    for prec in (_ffi.PREC_FP32, _ffi.PREC_TF32, _ffi.PREC_BF16):
        idx = _ffi.DeviceArray(ctx, (G, B), np.int32)
        best = _ffi.DeviceArray(ctx, (G, B), np.float32)
        gap = _ffi.DeviceArray(ctx, (G, B), np.float32)
        ctx.set_precision(prec)
        try:
            _ffi.check(L.pgmvae_vq_assign(ctx.h, None, dz.ptr, B * D, D, de.ptr, K * D, D, idx.ptr, B, best.ptr, gap.ptr,
                                          G, B, D, K))
        finally:
            ctx.set_precision(_ffi.PREC_FP32)
        out[prec] = (idx.numpy(), best.numpy(), gap.numpy())
        n = C.c_int(0)
        if prec != _ffi.PREC_FP32:
            _ffi.check(L.pgmvae_vq_assign_rescored(ctx.h, G, K, C.byref(n)))
        nres[prec] = n.value
    i32, b32, g32 = out[_ffi.PREC_FP32]
    idx_o, gap_o = O.vq_assign(z, np.ascontiguousarray(e.transpose(0, 2, 1)))
    safe = gap_o.numpy() > 1e-5
    for prec, nm in ((_ffi.PREC_TF32, "tf32"), (_ffi.PREC_BF16, "f16")):
        itc, btc, gtc = out[prec]
        mism = int((i32 != itc).sum())
        print(f"vq tc {nm}: G={G} B={B} D={D} K={K}: full-scan rows {nres[prec]}/{G * B}, mismatches {mism}")
        assert mism == 0
        assert 0 <= nres[prec] <= G * B // 100 + 1
        np.testing.assert_array_equal(btc, b32)          # candidates are re-scored with the fp32 arithmetic
        np.testing.assert_array_equal(itc[safe], idx_o.numpy()[safe])


def test_vq_assign_tensor_core_dead_codes(ctx):
    """Many identical (dead, all-zero) codes tie exactly: the candidate list overflows and the exact full
    scan must still return the lowest index, as tf.argmin does."""
    from pgmvae import _ffi
    G, B, D, K = 2, 700, 16, 256
    rng = np.random.default_rng(3)
    e = rng.uniform(-0.3, 0.3, (G, K, D)).astype(np.float32)
    e[:, 40:, :] = 0.0                                      # 216 dead codes
    z = (0.02 * rng.standard_normal((G, B, D))).astype(np.float32)      # closest to the zero vector
    dz, de = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, e)
    res = {}
    for prec in (_ffi.PREC_FP32, _ffi.PREC_TF32, _ffi.PREC_BF16):
        idx = _ffi.DeviceArray(ctx, (G, B), np.int32)
        ctx.set_precision(prec)
        try:
            _ffi.check(_ffi.lib().pgmvae_vq_assign(ctx.h, None, dz.ptr, B * D, D, de.ptr, K * D, D, idx.ptr, B, None, None,
                                                   G, B, D, K))
        finally:
            ctx.set_precision(_ffi.PREC_FP32)
        res[prec] = idx.numpy()
    np.testing.assert_array_equal(res[_ffi.PREC_TF32], res[_ffi.PREC_FP32])
    np.testing.assert_array_equal(res[_ffi.PREC_BF16], res[_ffi.PREC_FP32])
    assert (res[_ffi.PREC_FP32] == 40).mean() > 0.5


@pytest.mark.parametrize("G,B,D,K,sub", [(1, 3000, 64, 8192, 2), (1, 3000, 64, 8192, 3), (3, 700, 16, 200, 2),
                                         (2, 515, 30, 97, 3), (1, 100, 126, 300, 2)])
def test_vq_assign_ema_fused_equals_separate(ctx, G, B, D, K, sub, monkeypatch):
    """Fused single-pass fp16 tensor-core assignment + EMA scatter (pgmvae_vq_assign_ema) against the exact
    fp32 assignment followed by the stand-alone scatter: indices and counts bit-exact, sums to fp32 rounding."""
    from pgmvae import _ffi
    monkeypatch.setenv("PGMVAE_VQ_SUB", str(sub))
    rng = np.random.default_rng(G * 7 + B + D + K)
    e = (rng.uniform(-1, 1, (G, K, D)) * np.sqrt(3.0 / D)).astype(np.float32)
    pick = rng.integers(0, K, (G, B))
    z = (np.take_along_axis(e, pick[..., None], 1) + 0.3 * rng.standard_normal((G, B, D))).astype(np.float32)
    dz, de = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, e)
    L = _ffi.lib()
    idx_ref = _ffi.DeviceArray(ctx, (G, B), np.int32)
    cnt_ref, dw_ref = _ffi.DeviceArray(ctx, (G, K)), _ffi.DeviceArray(ctx, (G, K, D))
    _ffi.check(L.pgmvae_vq_assign(ctx.h, None, dz.ptr, B * D, D, de.ptr, K * D, D, idx_ref.ptr, B, None, None, G, B, D, K))
    _ffi.check(L.pgmvae_ema_stats(ctx.h, None, dz.ptr, B * D, D, idx_ref.ptr, B, cnt_ref.ptr, K, dw_ref.ptr, K * D, D,
                                  G, B, D, K))
    idx = _ffi.DeviceArray(ctx, (G, B), np.int32)
    cnt, dw = _ffi.DeviceArray(ctx, (G, K)), _ffi.DeviceArray(ctx, (G, K, D))
    _ffi.check(L.pgmvae_vq_assign_ema(ctx.h, None, dz.ptr, B * D, D, de.ptr, K * D, D, idx.ptr, B, cnt.ptr, K, dw.ptr,
                                      K * D, D, G, B, D, K))
    n = C.c_int(0)
    _ffi.check(L.pgmvae_vq_assign_rescored(ctx.h, G, K, C.byref(n)))
    print(f"fused vq+ema: G={G} B={B} D={D} K={K} sub={sub}: full-scan rows {n.value}/{G * B}")
    assert n.value <= G * B // 50 + 1
    np.testing.assert_array_equal(idx.numpy(), idx_ref.numpy())
    np.testing.assert_array_equal(cnt.numpy(), cnt_ref.numpy())
    assert cnt.numpy().sum() == G * B
    np.testing.assert_allclose(dw.numpy(), dw_ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("scale", [1e-2, 1e-3])
def test_vq_assign_tensor_core_small_magnitudes(ctx, scale):
    """Initialisation-like magnitudes (latents and codes ~1e-2): the candidate margin is relative to the
    magnitudes, so the tensor-core paths must neither lose exactness nor send every row to the full scan."""
    from pgmvae import _ffi
    G, B, D, K = 4, 2000, 64, 512
    rng = np.random.default_rng(17)
    e = (rng.uniform(-1, 1, (G, K, D)) * scale * 0.5).astype(np.float32)
    z = (rng.standard_normal((G, B, D)) * scale).astype(np.float32)
    dz, de = _ffi.DeviceArray.from_numpy(ctx, z), _ffi.DeviceArray.from_numpy(ctx, e)
    L = _ffi.lib()
    out = {}
    for prec in (_ffi.PREC_FP32, _ffi.PREC_TF32, _ffi.PREC_BF16):
        idx = _ffi.DeviceArray(ctx, (G, B), np.int32)
        ctx.set_precision(prec)
        try:
            _ffi.check(L.pgmvae_vq_assign(ctx.h, None, dz.ptr, B * D, D, de.ptr, K * D, D, idx.ptr, B, None, None, G, B, D, K))
        finally:
            ctx.set_precision(_ffi.PREC_FP32)
        n = C.c_int(0)
        if prec != _ffi.PREC_FP32:
            _ffi.check(L.pgmvae_vq_assign_rescored(ctx.h, G, K, C.byref(n)))
            assert n.value <= G * B // 50, (prec, n.value)
        out[prec] = idx.numpy()
    np.testing.assert_array_equal(out[_ffi.PREC_TF32], out[_ffi.PREC_FP32])
    np.testing.assert_array_equal(out[_ffi.PREC_BF16], out[_ffi.PREC_FP32])
