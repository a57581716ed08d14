"""GPU parity tests of the tcgen05 tensor-core path (ctx precision TF32): the grouped dense
GEMMs (forward / dgrad / wgrad with their fused epilogues) and the VQ assignment run on the
tensor cores; results are compared with the oracle and with the exact-fp32 CUDA-core path.
tf32 keeps 10 explicit mantissa bits (operands are truncated by the MMA), so the floating
point tolerances here are 10x looser than in test_gpu_model.py; codes stay exact because the
VQ kernel re-scores its candidates in fp32."""
import ctypes as C

import numpy as np
import pytest
import torch

import pgmvae_oracle as O
from test_oracle import load_case, make_oracle

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.fixture
def tf32(ctx):
    from pgmvae import _ffi
    ctx.set_precision(_ffi.PREC_TF32)
    yield ctx
    ctx.set_precision(_ffi.PREC_FP32)


@pytest.mark.parametrize("V,B,fin,fout,act", [
    (16, 256, 16, 12, "selu"), (3, 1, 4, 12, "selu"), (5, 200, 68, 52, "selu"), (2, 129, 132, 68, "sigmoid"),
    (4, 64, 16, 24, None), (69, 300, 20, 16, "selu"), (2, 1000, 400, 200, "selu"), (1, 333, 1556, 400, "selu"),
])
def test_fatdense_forward_tensor_core(tf32, V, B, fin, fout, act):
    from core.dense import FatDense
    rng = np.random.default_rng(V * 1000 + B)
    x = rng.standard_normal((V, B, fin)).astype(np.float32)
    layer = FatDense(fout, activation=act, kernel_initializer="he_uniform")
    layer.build(x.shape)
    layer.bias = rng.standard_normal((V, 1, fout)).astype(np.float32) * 0.1
    l0 = tf32.launches
    got = layer(x).numpy()
    exp = O.fatdense_call(torch.from_numpy(x), torch.from_numpy(layer.kernel), torch.from_numpy(layer.bias), act).numpy()
    err = np.abs(got - exp).max()
    scale = np.abs(x).max() * np.abs(layer.kernel).max() * np.sqrt(fin)
    print(f"dense tc fwd V={V} B={B} {fin}->{fout}: max abs err {err:.2e} (scale {scale:.2e})")
    assert err <= 4e-3 * scale


def _grads_case(name, prec_ctx):
    from core.model import VqVAE, Adam
    from pgmvae import _ffi
    z, cfg, params = load_case(name)
    m = VqVAE(cfg["units"], cfg["V"], cfg["D"], cfg["K"], cost=cfg["cost"], decay=cfg["decay"], ema=cfg["ema"],
              max_batch=max(cfg["B"], 256))
    m.set_weights_from(params)
    y = np.ascontiguousarray(z["y_train"][0])
    met = (C.c_double * 4)()
    _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y.ctypes.data, 0, y.shape[0], y.shape[0], cfg["lr"], None, 1, met))
    return z, cfg, m, list(met)


@pytest.mark.parametrize("name", ["v4_ema", "v9_grad", "v16_ema"])
def test_gradients_tensor_core(tf32, name):
    z, cfg, m, met = _grads_case(name, tf32)
    np.testing.assert_allclose(met[:3], z["metrics"][0][:3], rtol=2e-3)
    # tf32 GEMMs move z by ~1e-3, so samples whose two nearest codes are that close may switch code; such a
    # switch changes the decoder input of that sample and, in these tiny nets, a visible part of the gradient.
    idx = m(np.ascontiguousarray(z["y_train"][0]), code_only=True).argmax(-1)
    flips = int((idx != z["act.idx"]).sum())
    assert np.all(z["act.gap"][idx != z["act.idx"]] < 2e-2)
    tol = 3e-2 if flips == 0 else 0.5
    worst = 0.0
    for k in z.files:
        if k.startswith("grad1."):
            e = rel_err(m._get_tensor("grad." + k[6:]), z[k])
            worst = max(worst, e)
            assert e < tol, (k, e, flips)
    print(f"{name}: tf32 gradient max rel err {worst:.2e} ({flips} code switches)")


@pytest.mark.parametrize("ema,B", [(True, 256), (False, 300)])
def test_training_tensor_core_vs_oracle_plants_scale(tf32, ema, B):
    from core.model import VqVAE, Adam
    units, V, D, K = [50, 40, 30, 20], 69, 16, 128
    params = O.init_params(units, V, D, K, seed=9)
    cfg = dict(units=units, V=V, D=D, K=K, cost=0.25, decay=0.99, ema=ema, B=B, lr=1e-3)
    m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, max_batch=512)
    m.set_weights_from({k: v.numpy() for k, v in params.items()})
    m.compile(optimizer=Adam(lr=1e-3))
    om = make_oracle(cfg, {k: v.numpy() for k, v in params.items()})
    ys = O.synthetic_binary(3 * B, V, seed=2).reshape(3, B, V)
    for s in range(3):
        met = m.train_on_batch(np.ascontiguousarray(ys[s]))
        exp = om.train_step(O.make_xs(ys[s]), lr=1e-3)
        for k in ("loss", "mse", "mae"):
            assert abs(met[k] - exp[k]) <= 1e-3 * abs(exp[k]), (s, k, met[k], exp[k])
        assert abs(met["vq_loss"] - exp["vq_loss"]) <= 2e-2 * abs(exp["vq_loss"]) + 1e-9, (s, met, exp)
    n1, n0 = m.count(ys.reshape(-1, V))
    assert (n1 + n0).sum() == 3 * B * V


def test_tensor_core_matches_fp32_path_cfg2_full_batch(ctx):
    """Same weights, same batch (B=4096): tf32 tensor-core step vs exact-fp32 CUDA-core step."""
    from core.model import VqVAE, Adam
    from pgmvae import _ffi, data
    V, D, K, B = 69, 16, 128, 4096
    y = data.synthetic_binary(B, V, seed=3)
    res = {}
    for prec in (_ffi.PREC_FP32, _ffi.PREC_TF32):
        ctx.set_precision(prec)
        try:
            m = VqVAE([50, 40, 30, 20], V, D, K, cost=0.25, decay=0.99, ema=True, seed=1, max_batch=B)
            met = (C.c_double * 4)()
            _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y.ctypes.data, 0, B, B, 1e-3, None, 1, met))
            res[prec] = (list(met), {n: m._get_tensor("grad." + n) for n in ("fd0.kernel", "fd4.kernel", "fd9.kernel", "fd9.bias")},
                         m._get_tensor("vq.stat_c"))
        finally:
            ctx.set_precision(_ffi.PREC_FP32)
    a, b = res[_ffi.PREC_FP32], res[_ffi.PREC_TF32]
    np.testing.assert_allclose(b[0][:3], a[0][:3], rtol=1e-3)
    for n in a[1]:
        assert rel_err(b[1][n], a[1][n]) < 2e-2, n
    # a handful of samples may flip codes because z itself differs by tf32 rounding
    assert np.abs(a[2] - b[2]).sum() <= 0.01 * B * V


def _dev(ctx, a):
    from pgmvae import _ffi
    return _ffi.DeviceArray.from_numpy(ctx, np.ascontiguousarray(a))


@pytest.mark.parametrize("G,B,fin,fout", [(9, 33, 9, 8), (3, 7, 4, 6), (5, 64, 16, 15), (4, 300, 69, 50), (2, 4096, 50, 40),
                                          (2, 1000, 400, 200), (1, 513, 1556, 400)])
def test_dgrad_wgrad_operators_tensor_core_vs_fp32(ctx, G, B, fin, fout):
    """pgmvae_dense_dgrad / pgmvae_dense_wgrad on padded (multiple-of-8) layouts: tcgen05 vs CUDA cores."""
    from pgmvae import _ffi
    L = _ffi.lib()
    pin, pout = (fin + 7) // 8 * 8, (fout + 7) // 8 * 8
    rng = np.random.default_rng(G * 100 + B)
    x = np.zeros((G, B, pin), np.float32); x[..., :fin] = rng.standard_normal((G, B, fin))
    dy = np.zeros((G, B, pout), np.float32); dy[..., :fout] = rng.standard_normal((G, B, fout))
    w = np.zeros((G, pin, pout), np.float32); w[:, :fin, :fout] = rng.standard_normal((G, fin, fout)) * 0.3
    h = np.zeros((G, B, pin), np.float32); h[..., :fin] = rng.standard_normal((G, B, fin))
    dx_ref = (dy[..., :fout].astype(np.float64) @ w[:, :fin, :fout].transpose(0, 2, 1)) * np.where(
        h[..., :fin] < 0, h[..., :fin] + 1.7580993408473768, 1.0507009873554805)
    dw_ref = x[..., :fin].astype(np.float64).transpose(0, 2, 1) @ dy[..., :fout]
    db_ref = dy[..., :fout].astype(np.float64).sum(1)
    dX, dW, dDy, dH = _dev(ctx, x), _dev(ctx, w), _dev(ctx, dy), _dev(ctx, h)
    for prec, tol in ((_ffi.PREC_FP32, 2e-5), (_ffi.PREC_TF32, 3e-3)):
        ctx.set_precision(prec)
        try:
            dx = _ffi.DeviceArray(ctx, (G, B, pin)); dw = _ffi.DeviceArray(ctx, (G, pin, pout)); db = _ffi.DeviceArray(ctx, (G, pout))
            _ffi.check(L.pgmvae_dense_dgrad(ctx.h, None, dDy.ptr, B * pout, pout, dW.ptr, pin * pout, pout, dH.ptr, B * pin, pin,
                                            None, None, 0, 0, 0.0, dx.ptr, B * pin, pin, G, B, fin, fout, _ffi.ACT_SELU))
            _ffi.check(L.pgmvae_dense_wgrad(ctx.h, None, dX.ptr, B * pin, pin, dDy.ptr, B * pout, pout, dw.ptr, pin * pout, pout,
                                            db.ptr, pout, G, B, fin, fout, -1))
        finally:
            ctx.set_precision(_ffi.PREC_FP32)
        e1 = rel_err(dx.numpy()[..., :fin], dx_ref)
        e2 = rel_err(dw.numpy()[:, :fin, :fout], dw_ref)
        e3 = rel_err(db.numpy()[:, :fout], db_ref)
        print(f"G={G} B={B} {fin}x{fout} prec={prec}: dgrad {e1:.2e} wgrad {e2:.2e} db {e3:.2e}")
        # tensor-core path: the bias gradient is a ones-row MMA on the tf32 operands
        assert e1 < tol and e2 < tol and e3 < (1e-5 if prec == _ffi.PREC_FP32 else tol), (prec, e1, e2, e3)
        assert np.all(dx.numpy()[..., fin:] == 0) and np.all(dw.numpy()[:, fin:, :] == 0) and np.all(dw.numpy()[:, :, fout:] == 0)


@pytest.mark.parametrize("V,units,D,K,B,ema", [(16, [15, 14, 13, 12], 4, 32, 256, True), (69, [50, 40, 30, 20], 16, 128, 1000, True),
                                               (69, [50, 40, 30, 20], 16, 128, 515, False), (9, [70, 33, 9, 20], 8, 50, 77, True),
                                               (5, [6, 5, 4, 3], 3, 7, 1, True), (4, [120, 9, 64, 8], 30, 33, 385, False)])
@pytest.mark.parametrize("exact", [True, False])
def test_chain_kernels_equal_layer_by_layer_kernels(ctx, monkeypatch, V, units, D, K, B, ema, exact):
    """The TMEM-resident chain kernels (forward, backward, encode + histogram) against the layer-by-layer
    tcgen05 kernels on the same weights and batches: same tf32 products, so everything agrees to well
    within the 1e-3 bar; codes and PLL counts may differ only for latents on a decision boundary."""
    from core.model import VqVAE
    from pgmvae import _ffi, data
    y = data.synthetic_binary(2 * B, V, seed=5)
    names = [f"fd{l}.{t}" for l in range(10) for t in ("kernel", "bias")]
    res = {}
    # exact: the chains use expf like the layer kernels -> everything agrees to summation order.
    # default (ex2.approx activations): a few latents on a decision boundary flip their code, which moves the
    # gradients of that variable by O(1/B) per flip.
    if exact:
        monkeypatch.setenv("PGMVAE_CHAIN_EXACT", "1")
    gtol = 2e-4 if exact else 0.15
    ctx.set_precision(_ffi.PREC_TF32)
    try:
        for mode in ("layers", "chain"):
            if mode == "layers":
                monkeypatch.setenv("PGMVAE_NO_CHAIN", "1")
            else:
                monkeypatch.delenv("PGMVAE_NO_CHAIN", raising=False)
            m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, seed=2, max_batch=B)
            mets = []
            met = (C.c_double * 4)()
            # gradients of one step from identical weights (flag 1: no optimiser / EMA update)
            _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y[:B].ctypes.data, 0, B, B, 1e-3, None, 1, met))
            mets.append(list(met))
            grads = {n: m._get_tensor("grad." + n) for n in names}
            for s in range(2):
                _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y[s * B:(s + 1) * B].ctypes.data, 0, B, B, 1e-3, None,
                                                              0, met))
                mets.append(list(met))
            emb = m._get_tensor("vq.embeddings")
            n1, n0 = m.count(y)
            idx = m(y[:B], code_only=True).argmax(-1)
            res[mode] = (mets, grads, None, emb, n1, n0, idx)
    finally:
        ctx.set_precision(_ffi.PREC_FP32)
    a, b = res["layers"], res["chain"]
    # (the chains evaluate selu / sigmoid with ex2.approx, the layer kernels with expf)
    ma, mb = np.array(a[0]), np.array(b[0])
    np.testing.assert_allclose(mb[:, :3], ma[:, :3], rtol=2e-5 if exact else 1e-4)
    np.testing.assert_allclose(mb[:, 3], ma[:, 3], rtol=2e-2, atol=1e-9)       # vq_loss: a few codes may flip
    for n in names:
        assert rel_err(b[1][n], a[1][n]) < gtol, ("grad", n, rel_err(b[1][n], a[1][n]))
    assert rel_err(b[3], a[3]) < (1e-4 if exact else 1e-2)
    # codes: identical unless a latent lands within rounding of a decision boundary
    assert (a[6] != b[6]).mean() <= 2e-3
    assert (a[4] + a[5]).sum() == (b[4] + b[5]).sum() == 2 * B * V
    assert np.abs(a[4].astype(np.int64) - b[4].astype(np.int64)).sum() <= 4e-3 * B * V


def _step_snapshot(ctx, monkeypatch, env, V, units, D, K, B, ema=True, steps=2):
    """Gradients of one step and the state after `steps` optimiser steps with the given environment switches."""
    from core.model import VqVAE
    from pgmvae import _ffi, data
    for k in ("PGMVAE_WGRAD_PER_LAYER", "PGMVAE_CHAIN_SPLIT", "PGMVAE_CHAIN_NO_TAIL", "PGMVAE_NO_CHAIN"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    y = data.synthetic_binary(steps * B, V, seed=7)
    names = [f"fd{l}.{t}" for l in range(10) for t in ("kernel", "bias")]
    ctx.set_precision(_ffi.PREC_TF32)
    try:
        m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, seed=3, max_batch=B)
        met = (C.c_double * 4)()
        _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y[:B].ctypes.data, 0, B, B, 1e-3, None, 1, met))
        first = list(met)
        grads = {n: m._get_tensor("grad." + n) for n in names}
        mets = []
        for s in range(steps):
            _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y[s * B:(s + 1) * B].ctypes.data, 0, B, B, 1e-3, None, 0, met))
            mets.append(list(met))
        return first, grads, np.array(mets), m._get_tensor("vq.embeddings"), m._get_tensor("fd3.kernel")
    finally:
        ctx.set_precision(_ffi.PREC_FP32)


@pytest.mark.parametrize("switch", ["PGMVAE_WGRAD_PER_LAYER", "PGMVAE_CHAIN_SPLIT", "PGMVAE_CHAIN_NO_TAIL"])
@pytest.mark.parametrize("V,units,D,K,B,ema", [(69, [50, 40, 30, 20], 16, 128, 4096, True),
                                               (16, [15, 14, 13, 12], 4, 32, 53, False),
                                               (200, [20, 12, 9, 8], 8, 16, 300, True)])
def test_fast_paths_equal_their_unfused_counterparts(ctx, monkeypatch, switch, V, units, D, K, B, ema):
    """The default training step -- forward + dgrad stages in one chain launch, all weight-gradient GEMMs in one
    launch, left-over items as single tiles -- against the same step with one of these switched off: the arithmetic
    is the same (tf32 products, fp32 accumulation), only the order of the fp32 reductions over the batch differs."""
    a = _step_snapshot(ctx, monkeypatch, {}, V, units, D, K, B, ema)
    b = _step_snapshot(ctx, monkeypatch, {switch: "1"}, V, units, D, K, B, ema)
    np.testing.assert_allclose(b[0][:3], a[0][:3], rtol=1e-6)
    for n in a[1]:
        assert rel_err(b[1][n], a[1][n]) < 2e-5, (switch, n, rel_err(b[1][n], a[1][n]))
    np.testing.assert_allclose(b[2][:, :3], a[2][:, :3], rtol=1e-5)
    assert rel_err(b[3], a[3]) < 1e-5 and rel_err(b[4], a[4]) < 1e-4


def test_single_tile_tail_with_two_chains(ctx, monkeypatch):
    """Two chains per CTA (what a network with wider layers gets): 1104 items over 148 CTAs leave 68, which run
    as 136 single tiles; same results as whole left-over items."""
    V, units, D, K, B = 69, [50, 40, 30, 20], 16, 128, 4096
    monkeypatch.setenv("PGMVAE_CHAINS", "2")
    try:
        a = _step_snapshot(ctx, monkeypatch, {}, V, units, D, K, B, True, steps=1)
        b = _step_snapshot(ctx, monkeypatch, {"PGMVAE_CHAIN_NO_TAIL": "1"}, V, units, D, K, B, True, steps=1)
    finally:
        monkeypatch.delenv("PGMVAE_CHAINS", raising=False)
    np.testing.assert_allclose(b[0][:3], a[0][:3], rtol=1e-6)
    for n in a[1]:
        assert rel_err(b[1][n], a[1][n]) < 2e-5, (n, rel_err(b[1][n], a[1][n]))
    assert rel_err(b[3], a[3]) < 1e-5


def test_count_in_slabs_equals_count_in_batches(ctx, monkeypatch):
    """Stage 2 walks the data in slabs of up to 32768 samples when the chains are in use; the counts are the sums
    of the counts of the training-batch-sized pieces (exactly: integers)."""
    from core.model import VqVAE
    from pgmvae import _ffi, data
    V, units, D, K, B = 69, [50, 40, 30, 20], 16, 128, 512
    y = data.synthetic_binary(40000, V, seed=9)
    ctx.set_precision(_ffi.PREC_TF32)
    try:
        m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=True, seed=4, max_batch=B)
        n1, n0 = m.count(y)                                  # 32768 + 7232 rows
        p1 = np.zeros_like(n1)
        p0 = np.zeros_like(n0)
        for s in range(0, len(y), B):                        # <= max_batch rows per call: the per-batch path
            a1, a0 = m.count(y[s:s + B])
            p1 += a1
            p0 += a0
    finally:
        ctx.set_precision(_ffi.PREC_FP32)
    assert (n1 + n0).sum() == len(y) * V
    assert np.array_equal(n1, p1) and np.array_equal(n0, p0)
