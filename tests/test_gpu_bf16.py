"""GPU parity tests of the bf16 tcgen05 path (ctx precision BF16): the persistent grouped GEMMs of
csrc/dense_bf16.cu behind the operator ABI (pgmvae_dense_fwd / _fwd_sigmoid_mse / _dgrad / _wgrad,
reference core/dense.py:99-111 and its autodiff).

Two bars per operator:
  * against float64 numpy on the SAME bf16-rounded operands: what is left is fp32 accumulation order and the
    approximate activations (ex2.approx / rcp.approx), so the tolerance is tight (1e-4 of the output scale) --
    this pins tiling, descriptors, swizzles, ragged edges and the epilogues exactly;
  * against the fp32 oracle on the unrounded operands: the bf16 operand rounding itself (2^-9 per element),
    stated as a fraction of the output scale.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import pgmvae_oracle as O

pytestmark = pytest.mark.gpu

SELU_SCALE, SELU_SA = 1.0507009873554805, 1.7580993408473768


def bf16r(a):
    """round-to-nearest-even bf16 image of a float32 array, as float64"""
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).bfloat16().double().numpy()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.fixture
def bf16(ctx):
    from pgmvae import _ffi
    ctx.set_precision(_ffi.PREC_BF16)
    yield ctx
    ctx.set_precision(_ffi.PREC_FP32)


def _dev(ctx, a):
    from pgmvae import _ffi
    return _ffi.DeviceArray.from_numpy(ctx, np.ascontiguousarray(a))


def _act(x, act):
    if act == "selu":
        return np.where(x > 0, SELU_SCALE * x, SELU_SA * (np.exp(np.minimum(x, 0)) - 1.0))
    if act == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    return x


@pytest.mark.parametrize("V,B,fin,fout,act", [
    (16, 256, 16, 12, "selu"), (3, 1, 4, 12, "selu"), (5, 200, 68, 52, "selu"), (2, 129, 132, 68, "sigmoid"),
    (4, 64, 16, 24, None), (69, 300, 20, 16, "selu"), (2, 1000, 400, 200, "selu"), (1, 333, 1556, 400, "selu"),
    (3, 4096, 64, 50, "selu"), (2, 700, 400, 1556, "sigmoid"), (150, 130, 40, 30, "selu"),
])
@pytest.mark.parametrize("pair", [None, "1", "2", "3", "4"])
def test_fatdense_forward_bf16(bf16, monkeypatch, V, B, fin, fout, act, pair):
    from core.dense import FatDense
    if pair:
        monkeypatch.setenv("PGMVAE_BF16_PAIR", pair)
    rng = np.random.default_rng(V * 1000 + B)
    x = rng.standard_normal((V, B, fin)).astype(np.float32)
    layer = FatDense(fout, activation=act, kernel_initializer="he_uniform")
    layer.build(x.shape)
    layer.bias = rng.standard_normal((V, 1, fout)).astype(np.float32) * 0.1
    got = layer(x).numpy()
    pre = bf16r(x) @ bf16r(layer.kernel) + layer.bias.astype(np.float64)
    exact = _act(pre, act)
    scale = max(np.abs(exact).max(), 1e-6)
    e_same = np.abs(got - exact).max() / scale
    exp = O.fatdense_call(torch.from_numpy(x), torch.from_numpy(layer.kernel), torch.from_numpy(layer.bias), act).numpy()
    e_oracle = np.abs(got - exp).max() / scale
    print(f"dense bf16 fwd V={V} B={B} {fin}->{fout}: vs bf16-rounded operands {e_same:.2e}, vs fp32 oracle {e_oracle:.2e}")
    assert e_same <= 1e-4
    assert e_oracle <= 2e-2


@pytest.mark.parametrize("G,B,fin,V,g0", [(3, 50, 12, 16, 2), (2, 300, 50, 69, 10), (2, 513, 400, 1556, 700), (5, 128, 20, 40, 35)])
@pytest.mark.parametrize("pair", [None, "1", "2", "3", "4"])
def test_fwd_sigmoid_mse_bf16(bf16, monkeypatch, G, B, fin, V, g0, pair):
    """fd9 + loss: sigmoid output, squared / absolute error sums with the leave-one-out column masked, and
    d(loss)/d(pre-activation) (stored in bf16 by the kernel)."""
    from pgmvae import _ffi
    L = _ffi.lib()
    if pair:
        monkeypatch.setenv("PGMVAE_BF16_PAIR", pair)
    pin, pv = (fin + 7) // 8 * 8, (V + 7) // 8 * 8
    rng = np.random.default_rng(G * 10 + B)
    x = np.zeros((G, B, pin), np.float32); x[..., :fin] = rng.standard_normal((G, B, fin))
    w = np.zeros((G, pin, pv), np.float32); w[:, :fin, :V] = rng.standard_normal((G, fin, V)) * (1.0 / np.sqrt(fin))
    b = np.zeros((G, pv), np.float32); b[:, :V] = rng.standard_normal((G, V)) * 0.1
    y = np.zeros((B, pv), np.float32); y[:, :V] = rng.random((B, V)) < 0.3
    gscale = 2.0 / (B * V * (V - 1))
    dX, dW, dB, dY = _dev(bf16, x), _dev(bf16, w), _dev(bf16, b), _dev(bf16, y)
    dpre = _ffi.DeviceArray(bf16, (G, B, pv)); out = _ffi.DeviceArray(bf16, (G, B, pv)); acc = _ffi.DeviceArray(bf16, (4,), np.float64)
    _ffi.check(L.pgmvae_dense_fwd_sigmoid_mse(bf16.h, None, dX.ptr, B * pin, pin, dW.ptr, pin * pv, pv, dB.ptr, pv, dY.ptr, pv,
                                              dpre.ptr, B * pv, pv, out.ptr, acc.ptr, G, g0, B, fin, V, C.c_float(gscale)))
    o = 1.0 / (1.0 + np.exp(-(bf16r(x[..., :fin]) @ bf16r(w[:, :fin, :V]) + b[:, None, :V].astype(np.float64))))
    d = o - y[None, :, :V]
    for g in range(G):
        d[g, :, g0 + g] = 0.0
    dp = gscale * d * o * (1.0 - o)
    a = acc.numpy()
    e_out = rel_err(out.numpy()[..., :V], o)
    e_dp = np.abs(dpre.numpy()[..., :V] - dp).max() / np.abs(dp).max()
    print(f"fd9 bf16 G={G} B={B} {fin}->{V}: out {e_out:.2e} dpre {e_dp:.2e} sq {a[0] / (d * d).sum() - 1:.2e} ab {a[1] / np.abs(d).sum() - 1:.2e}")
    assert e_out <= 1e-4
    assert e_dp <= 5e-3                       # dpre is rounded to bf16 on the way out (2^-9)
    assert abs(a[0] / (d * d).sum() - 1) <= 1e-4 and abs(a[1] / np.abs(d).sum() - 1) <= 1e-4
    assert np.all(dpre.numpy()[np.arange(G), :, g0 + np.arange(G)] == 0)


@pytest.mark.parametrize("G,B,fin,fout", [(9, 33, 9, 8), (3, 7, 4, 6), (5, 64, 16, 15), (4, 300, 69, 50), (2, 4096, 50, 40),
                                          (2, 1000, 400, 200), (1, 513, 1556, 400), (2, 600, 400, 1556), (3, 256, 64, 50)])
@pytest.mark.parametrize("orient,pair", [("auto", None), ("d", None), ("t", None), ("d", "1"), ("d", "2"), ("t", "1"), ("t", "2"), ("d", "3"), ("t", "3"), ("d", "4"), ("t", "4")])
def test_dgrad_wgrad_operators_bf16(bf16, monkeypatch, G, B, fin, fout, orient, pair):
    """pgmvae_dense_dgrad / pgmvae_dense_wgrad on padded (multiple-of-8) layouts, both wgrad orientations, and with the
    2-CTA cluster schedule forced (pair 1: neighbouring M tiles share a multicast B tile; pair 2: neighbouring N tiles
    share A; pair 3: one cta_group::2 MMA over both SMs of the pair; pair 4: clusters of 2 x 2 tiles with both operands
    multicast; odd tile counts leave phantom tiles)."""
    from pgmvae import _ffi
    L = _ffi.lib()
    if orient != "auto":
        monkeypatch.setenv("PGMVAE_WGRAD_ORIENT", orient)
    if pair:
        monkeypatch.setenv("PGMVAE_BF16_PAIR", pair)
    pin, pout = (fin + 7) // 8 * 8, (fout + 7) // 8 * 8
    rng = np.random.default_rng(G * 100 + B)
    x = np.zeros((G, B, pin), np.float32); x[..., :fin] = rng.standard_normal((G, B, fin))
    dy = np.zeros((G, B, pout), np.float32); dy[..., :fout] = rng.standard_normal((G, B, fout))
    w = np.zeros((G, pin, pout), np.float32); w[:, :fin, :fout] = rng.standard_normal((G, fin, fout)) * 0.3
    h = np.zeros((G, B, pin), np.float32); h[..., :fin] = rng.standard_normal((G, B, fin))
    dsel = np.where(h[..., :fin] < 0, h[..., :fin].astype(np.float64) + SELU_SA, SELU_SCALE)
    dx_ref = (bf16r(dy[..., :fout]) @ bf16r(w[:, :fin, :fout]).transpose(0, 2, 1)) * dsel
    dw_ref = bf16r(x[..., :fin]).transpose(0, 2, 1) @ bf16r(dy[..., :fout])
    db_ref = bf16r(dy[..., :fout]).sum(1)
    dX, dW, dDy, dH = _dev(bf16, x), _dev(bf16, w), _dev(bf16, dy), _dev(bf16, h)
    dx = _ffi.DeviceArray(bf16, (G, B, pin)); dw = _ffi.DeviceArray(bf16, (G, pin, pout)); db = _ffi.DeviceArray(bf16, (G, pout))
    _ffi.check(L.pgmvae_dense_dgrad(bf16.h, None, dDy.ptr, B * pout, pout, dW.ptr, pin * pout, pout, dH.ptr, B * pin, pin,
                                    None, None, 0, 0, C.c_float(0.0), dx.ptr, B * pin, pin, G, B, fin, fout, _ffi.ACT_SELU))
    zero_row = 1 if fin > G + 1 else -1
    for _ in range(2):          # the operator accumulates: two calls = twice the gradient
        _ffi.check(L.pgmvae_dense_wgrad(bf16.h, None, dX.ptr, B * pin, pin, dDy.ptr, B * pout, pout, dw.ptr, pin * pout, pout,
                                        db.ptr, pout, G, B, fin, fout, zero_row))
    if zero_row >= 0:
        for g in range(G):
            dw_ref[g, zero_row + g, :] = 0.0
    e1 = rel_err(dx.numpy()[..., :fin], dx_ref)
    e2 = rel_err(dw.numpy()[:, :fin, :fout], 2 * dw_ref)
    e3 = rel_err(db.numpy()[:, :fout], 2 * db_ref)
    print(f"bf16 G={G} B={B} {fin}x{fout} orient={orient}: dgrad {e1:.2e} wgrad {e2:.2e} db {e3:.2e}")
    assert e1 < 1e-4 and e2 < 1e-4 and e3 < 1e-4, (e1, e2, e3)
    assert np.all(dx.numpy()[..., fin:] == 0) and np.all(dw.numpy()[:, fin:, :] == 0) and np.all(dw.numpy()[:, :, fout:] == 0)


def test_dgrad_commitment_gradient_bf16(bf16):
    """dgrad at the VQ boundary: + cscale (z - q) before act' (core/quantizer.py:142,153 through autodiff)."""
    from pgmvae import _ffi
    L = _ffi.lib()
    G, B, fin, fout = 4, 200, 16, 24
    rng = np.random.default_rng(5)
    dy = rng.standard_normal((G, B, fout)).astype(np.float32)
    w = (rng.standard_normal((G, fin, fout)) * 0.3).astype(np.float32)
    z = rng.standard_normal((G, B, fin)).astype(np.float32)
    q = rng.standard_normal((G, B, fin)).astype(np.float32)
    cs = 0.37
    dsel = np.where(z < 0, z.astype(np.float64) + SELU_SA, SELU_SCALE)
    ref = (bf16r(dy) @ bf16r(w).transpose(0, 2, 1) + cs * (z.astype(np.float64) - q)) * dsel
    dDy, dW, dZ, dQ = _dev(bf16, dy), _dev(bf16, w), _dev(bf16, z), _dev(bf16, q)
    dx = _ffi.DeviceArray(bf16, (G, B, fin))
    _ffi.check(L.pgmvae_dense_dgrad(bf16.h, None, dDy.ptr, B * fout, fout, dW.ptr, fin * fout, fout, dZ.ptr, B * fin, fin,
                                    dZ.ptr, dQ.ptr, B * fin, fin, C.c_float(cs), dx.ptr, B * fin, fin, G, B, fin, fout,
                                    _ffi.ACT_SELU))
    assert rel_err(dx.numpy(), ref) < 1e-4


# ---------------------------------------------------------------------------------------------- model level
def _model(ctx, prec, units, V, D, K, B, ema, params, monkeypatch, group_vars=None, no_chain=True):
    from core.model import VqVAE, Adam
    if no_chain:
        monkeypatch.setenv("PGMVAE_NO_CHAIN", "1")
    if group_vars:
        monkeypatch.setenv("PGMVAE_GROUP_VARS", str(group_vars))
    else:
        monkeypatch.delenv("PGMVAE_GROUP_VARS", raising=False)
    ctx.set_precision(prec)
    try:
        m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, max_batch=B)
    finally:
        from pgmvae import _ffi
        ctx.set_precision(_ffi.PREC_FP32)
    m.set_weights_from(params)
    m.compile(optimizer=Adam(lr=1e-3))
    return m


def _grads(m, y, lr=1e-3):
    from pgmvae import _ffi
    met = (C.c_double * 4)()
    _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y.ctypes.data, 0, y.shape[0], y.shape[0], lr, None, 1, met))
    names = [f"fd{l}.{t}" for l in range(10) for t in ("kernel", "bias")]
    return list(met), {n: m._get_tensor("grad." + n) for n in names}


GEOMS = [
    # V, units, D, K, B, ema, group_vars
    (16, [15, 14, 13, 12], 4, 32, 256, True, None),
    (69, [50, 40, 30, 20], 16, 128, 300, True, 20),          # 4 groups, the last one partial (9 variables)
    (24, [400, 200, 100, 50], 64, 512, 200, True, 7),        # cfg3-like widths, D = 64, K = 512, 4 groups
    (24, [400, 200, 100, 50], 64, 512, 130, False, 5),       # gradient-trained codebook
]


@pytest.mark.parametrize("pair", [None, "1", "2", "3", "4"])
@pytest.mark.parametrize("V,units,D,K,B,ema,gv", GEOMS)
def test_bf16_model_gradients_vs_oracle(ctx, monkeypatch, V, units, D, K, B, ema, gv, pair):
    """One step of the bf16 model (multi-group where gv is set) against the fp32 oracle: losses at 1e-3 / 2e-3,
    every gradient tensor at 3e-2 of its largest element (bf16 operands: 2^-9 per element; a code that flips
    because z moved by the rounding changes that sample's decoder input)."""
    from pgmvae import _ffi
    from test_oracle import make_oracle
    if pair:
        monkeypatch.setenv("PGMVAE_BF16_PAIR", pair)
    params = {k: v.numpy() for k, v in O.init_params(units, V, D, K, seed=11).items()}
    y = O.synthetic_binary(B, V, seed=4)
    m = _model(ctx, _ffi.PREC_BF16, units, V, D, K, B, ema, params, monkeypatch, gv)
    assert m.group_size() == (gv or V)
    met, grads = _grads(m, y)
    cfg = dict(units=units, V=V, D=D, K=K, cost=0.25, decay=0.99, ema=ema, B=B, lr=1e-3)
    om = make_oracle(cfg, params)
    exp, ograds = om.loss_and_grads(O.make_xs(y))
    exp["grads"] = {k: v.numpy() for k, v in ograds.items()}
    idx = m(y, code_only=True).argmax(-1)
    flips = float((idx != om.last_idx.numpy()).mean())
    print(f"bf16 model V={V} units={units}: loss {met[0]:.6f} vs {exp['loss']:.6f}, mse {met[1]:.6f} vs {exp['mse']:.6f}, "
          f"vq {met[3]:.3e} vs {exp['vq_loss']:.3e}, code flips {flips:.2e}")
    assert abs(met[1] - exp["mse"]) <= 1e-3 * exp["mse"]
    assert abs(met[2] - exp["mae"]) <= 1e-3 * exp["mae"]
    assert abs(met[3] - exp["vq_loss"]) <= 2e-2 * exp["vq_loss"] + 1e-9
    worst = worst_fro = 0.0
    for n, g in grads.items():
        e = rel_err(g, exp["grads"][n])
        ref = exp["grads"][n].astype(np.float64)
        fro = np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-30)
        worst, worst_fro = max(worst, e), max(worst_fro, fro)
        # element-wise maximum: a flipped code changes a whole sample's contribution; norm-wise: bf16 rounding
        assert e < (5e-2 if flips == 0 else 0.5), (n, e, flips)
        assert fro < (6e-2 if flips == 0 else 0.15), (n, fro, flips)
    print(f"   worst gradient error: max-norm {worst:.2e}, Frobenius {worst_fro:.2e}")


@pytest.mark.parametrize("V,units,D,K,B,ema,gv", GEOMS[1:])
@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
def test_multi_group_equals_single_group(ctx, monkeypatch, prec, V, units, D, K, B, ema, gv):
    """Walking the variables in groups (workspace smaller than the model: Vg < V) gives the same step as one
    group: the networks are independent, only losses / statistics / gradients are accumulated across groups."""
    from pgmvae import _ffi
    P = {"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32, "bf16": _ffi.PREC_BF16}[prec]
    params = {k: v.numpy() for k, v in O.init_params(units, V, D, K, seed=12).items()}
    y = O.synthetic_binary(2 * B, V, seed=5)
    res = {}
    for mode, g in (("one", None), ("many", gv)):
        m = _model(ctx, P, units, V, D, K, B, ema, params, monkeypatch, g)
        assert m.group_size() == (g or V)
        met, grads = _grads(m, y[:B])
        mets = [m.train_on_batch(np.ascontiguousarray(y[:B]))]
        emb1 = m._get_tensor("vq.embeddings")
        mets.append(m.train_on_batch(np.ascontiguousarray(y[B:])))
        n1, n0 = m.count(y)
        res[mode] = (met, grads, mets, m._get_tensor("vq.embeddings"), m._get_tensor("fd0.kernel"), n1, n0, emb1)
    a, b = res["one"], res["many"]
    np.testing.assert_allclose(b[0], a[0], rtol=1e-6)
    for n in a[1]:
        assert rel_err(b[1][n], a[1][n]) < 1e-6, (n, rel_err(b[1][n], a[1][n]))
    for s in range(2):
        for k in a[2][s]:
            assert abs(a[2][s][k] - b[2][s][k]) <= 1e-6 * abs(a[2][s][k]) + 1e-12
    # the codebook after ONE update is the same to rounding (the statistics are sums of fp32 reductions whose order
    # is not fixed); after the second step a latent within that rounding of a decision boundary may take the other
    # code, which moves one code vector by ~decay/size of it
    assert rel_err(b[7], a[7]) < 2e-6 and rel_err(b[4], a[4]) < 1e-5
    assert rel_err(b[3], a[3]) < 1e-2
    assert (a[5] + a[6]).sum() == (b[5] + b[6]).sum() == 2 * B * V
    assert np.abs(a[5] - b[5]).sum() <= 1e-3 * B * V


@pytest.mark.parametrize("prec,tol_loss", [("tf32", 1e-3), ("bf16", 1e-3)])
def test_three_steps_vs_oracle_cfg2_shapes_state_and_pll(ctx, monkeypatch, prec, tol_loss):
    """north_star's bar on the tensor-core paths at the cfg2 shapes: after three optimiser steps the losses, the EMA
    codebook, the CPT and the PLL agree with the fp32 oracle within 1e-3 relative; codes may differ only where the
    two nearest codes are closer than the rounding moved z (the flip fraction is reported)."""
    from pgmvae import _ffi
    from test_oracle import make_oracle
    units, V, D, K, B = [50, 40, 30, 20], 69, 16, 128, 512
    P = {"tf32": _ffi.PREC_TF32, "bf16": _ffi.PREC_BF16}[prec]
    params = {k: v.numpy() for k, v in O.init_params(units, V, D, K, seed=9).items()}
    m = _model(ctx, P, units, V, D, K, B, True, params, monkeypatch, None, no_chain=(prec == "bf16"))
    cfg = dict(units=units, V=V, D=D, K=K, cost=0.25, decay=0.99, ema=True, B=B, lr=1e-3)
    om = make_oracle(cfg, params)
    ys = O.synthetic_binary(3 * B, V, seed=2).reshape(3, B, V)
    for s in range(3):
        met = m.train_on_batch(np.ascontiguousarray(ys[s]))
        exp = om.train_step(O.make_xs(ys[s]), lr=1e-3)
        for k in ("loss", "mse", "mae"):
            assert abs(met[k] - exp[k]) <= tol_loss * abs(exp[k]), (s, k, met[k], exp[k])
        print(f"{prec} step {s}: vq_loss {met['vq_loss']:.4e} vs {exp['vq_loss']:.4e} "
              f"({abs(met['vq_loss'] - exp['vq_loss']) / exp['vq_loss']:.1e})")
    yall = ys.reshape(-1, V)
    emb, oemb = m._get_tensor("vq.embeddings"), om.p["vq.embeddings"].detach().numpy()
    used = np.abs(oemb).max(axis=1, keepdims=True) > 0                     # dead codes are exactly 0 on both sides
    e_emb = np.abs(emb - oemb).max() / np.abs(oemb).max()
    e_fro = np.linalg.norm(emb.astype(np.float64) - oemb) / np.linalg.norm(oemb.astype(np.float64))
    idx = m(yall, code_only=True).argmax(-1)
    with torch.no_grad():
        oidx = om(O.make_xs(yall), code_only=True).argmax(-1).numpy()
    flips = float((idx != oidx).mean())
    m.dist = m.cpt(yall)
    om.dist = om.cpt(O.make_xs(yall), yall)
    e_dist = np.abs(m.dist - om.dist.numpy()).max()
    pll, opll = m.pseudo_log_likelihood(yall), om.pseudo_log_likelihood(O.make_xs(yall), yall)
    print(f"{prec} after 3 steps: codebook max err {e_emb:.2e} of max |e| (Frobenius {e_fro:.2e}), code flips {flips:.2e}, "
          f"dist max abs err {e_dist:.2e}, pll {pll:.6f} vs {opll:.6f} ({abs(pll - opll) / abs(opll):.1e}); used codes {used.mean():.2f}")
    assert abs(pll - opll) <= 1e-3 * abs(opll)
    # a code vector is the mean of the few latents assigned to it, so its error is the rounding error of z itself
    # (2^-11 per tf32 operand, 2^-9 per bf16 operand, averaged over the contraction): the codebook as a whole agrees to
    # 1e-3 (tf32) / 3e-3 (bf16) in norm; single entries of rarely used codes deviate a few times more
    # This is the regime right after initialisation: the latents are tiny, ~1 % of the codes are alive and the alive
    # codes of a variable nearly coincide, so a latent moved by operand rounding may take the neighbouring code (the
    # flip fraction printed above).  Codes with a solid membership in the oracle (EMA cluster size >= 8) must agree;
    # a code that has a single member on one side and none on the other differs by its whole vector.
    size = om.ema_state.ema_cluster_size.numpy()                                     # [V, K]
    solid = (size >= 8.0)[:, None, :] & np.ones_like(oemb, dtype=bool)
    e_solid = np.linalg.norm((emb - oemb)[solid]) / max(np.linalg.norm(oemb[solid]), 1e-30)
    print(f"   codes with >= 8 members: {int((size >= 8.0).sum())}, their Frobenius error {e_solid:.2e}")
    if prec == "tf32":
        assert e_fro <= 1e-3 and e_emb <= 5e-3 and e_dist <= 2e-3 and flips <= 1e-3
    else:
        assert e_solid <= 5e-2 and flips <= 3e-2
