"""GPU parity tests, model level: the fused training step, VqVAE.count / cpt / PLL and the
run.py flow through the reference-facing API, against the committed oracle fixtures and the
live oracle.  Floating point within 1e-3 relative (north_star); counts and codes exact."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import pgmvae_oracle as O
from test_oracle import load_case, make_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["v4_ema", "v9_grad", "v16_ema"]


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def build(cfg, params, max_batch=None):
    from core.model import VqVAE, Adam
    m = VqVAE(cfg["units"], cfg["V"], cfg["D"], cfg["K"], cost=cfg["cost"], decay=cfg["decay"], ema=cfg["ema"],
              max_batch=max_batch or max(cfg["B"], 256))
    m.set_weights_from(params)
    m.compile(optimizer=Adam(lr=cfg["lr"]), loss="mse", metrics=["mae"])
    return m


@pytest.mark.parametrize("name", CASES)
def test_forward_activations_and_codes(ctx, name):
    z, cfg, params = load_case(name)
    m = build(cfg, params)
    y = z["y_train"][0]
    onehot = m(y, code_only=True)
    idx = onehot.argmax(-1)
    safe = z["act.gap"] > 1e-5
    np.testing.assert_array_equal(idx[safe], z["act.idx"][safe])
    rec = m(O.make_xs(y).numpy(), training=False)           # materialised reference input accepted
    assert rec.shape == (cfg["B"], cfg["V"], cfg["V"] - 1)
    assert rel_err(rec, z["act.out"]) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_gradients_first_step(ctx, name):
    import ctypes as C
    from pgmvae import _ffi
    z, cfg, params = load_case(name)
    m = build(cfg, params)
    y = np.ascontiguousarray(z["y_train"][0])
    met = (C.c_double * 4)()
    _ffi.check(_ffi.lib().pgmvae_model_train_step(m._h, y.ctypes.data, 0, y.shape[0], y.shape[0], cfg["lr"], None, 1, met))
    np.testing.assert_allclose(list(met), z["metrics"][0], rtol=1e-5)
    for k in z.files:
        if k.startswith("grad1."):
            got = m._get_tensor("grad." + k[6:])
            assert rel_err(got, z[k]) < 1e-4, k
    if cfg["ema"]:
        np.testing.assert_array_equal(m._get_tensor("vq.stat_c"), z["stat1.counts"])
        assert rel_err(m._get_tensor("vq.stat_w"), z["stat1.dw"]) < 1e-5


@pytest.mark.parametrize("name", CASES)
def test_three_training_steps_match_golden(ctx, name):
    z, cfg, params = load_case(name)
    m = build(cfg, params)
    for s in range(cfg["steps"]):
        met = m.train_on_batch(np.ascontiguousarray(z["y_train"][s]))
        exp = z["metrics"][s]
        np.testing.assert_allclose([met["loss"], met["mse"], met["mae"], met["vq_loss"]], exp, rtol=1e-3)
        if s in (0, cfg["steps"] - 1):
            tag = f"state{s + 1}."
            for k in z.files:
                if k.startswith(tag):
                    n = k[len(tag):]
                    assert rel_err(m._get_tensor(n), z[k]) < 1e-3, (s, n)


@pytest.mark.parametrize("name", CASES)
def test_count_cpt_pll_match_golden(ctx, name):
    z, cfg, params = load_case(name)
    m = build(cfg, params)
    last = f"state{cfg['steps']}."
    m.load_state_dict({k[len(last):]: z[k] for k in z.files if k.startswith(last) and "adam" not in k})
    y = z["y_eval"]
    # oracle codes and gaps for the evaluation split, to exclude numerically tied rows
    om = make_oracle(cfg, {k[len(last):]: z[k] for k in z.files if k.startswith(last) and ".adam_" not in k
                           and not k.endswith(("ema_w", "ema_cluster_size", "biased_w", "biased_c"))})
    keep = {}
    om(O.make_xs(y), code_only=True, keep=keep)
    _, gap = O.vq_assign(keep["h5"], om.p["vq.embeddings"])
    n1, n0 = m.count(y)
    if bool((gap > 1e-5).all()):
        np.testing.assert_array_equal(n1, z["n1"])
        np.testing.assert_array_equal(n0, z["n0"])
    assert np.abs(n1 - z["n1"]).sum() <= 2 * int((gap <= 1e-5).sum())
    assert (n1 + n0).sum() == y.shape[0] * cfg["V"]
    m.dist = m.cpt(O.make_xs(y).numpy(), y)                  # reference call shape cpt(x, y)
    assert rel_err(m.dist, z["dist"]) < 1e-3
    pll = m.pseudo_log_likelihood(y)
    assert abs(pll - float(z["pll"])) <= 1e-3 * abs(float(z["pll"]))


@pytest.mark.parametrize("ema,B", [(True, 256), (False, 53), (True, 1)])
def test_live_oracle_parity_plants_scale(ctx, ema, B):
    """cfg2 shapes (V=69, units 50/40/30/20, D=16, K=128) at a batch the oracle finishes quickly."""
    units, V, D, K = [50, 40, 30, 20], 69, 16, 128
    params = O.init_params(units, V, D, K, seed=9)
    cfg = dict(units=units, V=V, D=D, K=K, cost=0.25, decay=0.99, ema=ema, B=B, lr=1e-3)
    m = build(cfg, {k: v.numpy() for k, v in params.items()}, max_batch=256)
    om = make_oracle(cfg, {k: v.numpy() for k, v in params.items()})
    ys = O.synthetic_binary(2 * B, V, seed=2).reshape(2, B, V)
    for s in range(2):
        met = m.train_on_batch(np.ascontiguousarray(ys[s]))
        exp = om.train_step(O.make_xs(ys[s]), lr=1e-3)
        for k in ("loss", "mse", "mae", "vq_loss"):
            assert abs(met[k] - exp[k]) <= 1e-3 * abs(exp[k]) + 1e-9, (s, k, met[k], exp[k])
    st = om.state_numpy()
    # Adam's first steps move every weight by ~lr * g / (|g| + 3e-6): elements whose gradient is near that
    # epsilon amplify fp32 summation-order differences, so allow 1% of one lr-step on top of 1e-3 relative.
    # (B=1 is the ragged-size case: single-sample gradients sit at that epsilon, metrics above are the check.)
    for n in ["fd0.kernel", "fd4.bias", "fd5.kernel", "fd9.kernel", "fd9.bias", "vq.embeddings"] if B > 1 else []:
        got, ref = m._get_tensor(n), st[n]
        assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max() + 0.01 * 1e-3, n


def test_fit_with_pinned_order_and_partial_batch(ctx):
    z, cfg, params = load_case("v4_ema")
    m = build(cfg, params)
    om = make_oracle(cfg, params)
    y = z["y_eval"][:50]
    order = [np.random.default_rng(e).permutation(50) for e in range(2)]
    h = m.fit(y, y, batch_size=16, epochs=2, order=order)          # 3 full batches + one of 2
    ho = om.fit(O.make_xs(y), 16, 2, lr=cfg["lr"], order=order)
    w = np.array([16, 16, 16, 2] * 2, dtype=np.float64)
    lo = np.array([d["loss"] for d in ho])
    for e in range(2):
        exp = (lo[4 * e:4 * e + 4] * w[:4]).sum() / 50
        assert abs(h.history["loss"][e] - exp) <= 1e-3 * abs(exp)
    assert rel_err(m.fd3.kernel, om.p["fd3.kernel"].detach().numpy()) < 1e-3


def test_full_size_properties_cfg2(ctx):
    """BASELINE cfg2 at full size (B=4096): size-independent invariants."""
    from core.model import VqVAE
    from pgmvae import data
    V, D, K, B = 69, 16, 128, 4096
    m = VqVAE([50, 40, 30, 20], V, D, K, cost=0.25, decay=0.99, ema=True, seed=0, max_batch=B)
    y = data.synthetic_binary(3 * B, V, seed=0)
    losses = []
    for s in range(3):
        met = m.train_on_batch(np.ascontiguousarray(y[s * B:(s + 1) * B]))
        assert all(np.isfinite(v) for v in met.values())
        losses.append(met["loss"])
        # step-1 identity of the zero-debiased EMA: the visible counts are the raw histogram
        if s == 0:
            np.testing.assert_allclose(m.vq_layer.ema_cluster_size.sum(1), B, rtol=1e-5)
    assert 0.15 < losses[0] < 0.35                                   # sigmoid(0)=0.5 against 0/1 data
    n1, n0 = m.count(y)
    assert (n1 + n0).sum() == y.shape[0] * V
    np.testing.assert_array_equal((n1 + n0).sum(1), y.shape[0])
    np.testing.assert_array_equal(n1.sum(1), y.sum(0))               # ones per variable
    n1b, n0b = m.count(y)                                            # idempotent
    np.testing.assert_array_equal(n1, n1b)
    # counting in two halves adds up (linearity over samples)
    a1, a0 = m.count(y[:5000])
    b1, b0 = m.count(y[5000:])
    np.testing.assert_array_equal(a1 + b1, n1)
    m.dist = m.cpt(y)
    pll = m.pseudo_log_likelihood(y)
    assert -V * np.log(2) * 1.2 < pll < 0


def test_run_py_nltcs_end_to_end(ctx, tmp_path):
    """cfg1 flow: run.py on nltcs (few epochs), identifier and result line as the reference."""
    env = dict(os.environ, PGMVAE_DATA="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "pgm-vae_b200", "run.py"), "-n", "nltcs", "-k", "32", "-d", "4",
                        "-b", "256", "-e", "3", "--ema", "-u", "0"], capture_output=True, text=True, cwd=tmp_path, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = open(tmp_path / "result.txt").read().strip()
    assert line.startswith("nltcs_K-32_D-4_bs-256_epk-3_lr-0.001_bta-0.25_ema-True_gma-0.99_sd-0- pll-train:")
    assert line.endswith("cmll-test:1")
    vals = {k: float(v) for k, v in (t.split(":") for t in line.split(" ")[1:])}
    for k in ("pll-train", "pll-valid", "pll-test"):
        assert -16 * np.log(2) < vals[k] < -3.0


def test_gibbs_cmll_matches_oracle(ctx):
    """conditional_marginal_log_likelihood (reference core/model.py:110-148) with the same injected uniform draws:
    the device sub-net path (get_probability) against the oracle."""
    from core.model import VqVAE
    units, V, D, K, N = [15, 14, 13, 12], 16, 4, 32, 512
    params = {k: v.numpy() for k, v in O.init_params(units, V, D, K, seed=3).items()}
    y = O.synthetic_binary(N, V, seed=4)
    m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=True, max_batch=N)
    m.set_weights_from(params)
    om = O.OracleVqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=True,
                       params={k: torch.from_numpy(v) for k, v in params.items()})
    m.dist = m.cpt(y)
    om.dist = om.cpt(O.make_xs(y), y)
    xt = y[:24].astype(np.float32)
    draws = np.random.default_rng(9).random((200, 6, 24), dtype=np.float32)
    def source():
        it = iter(draws)
        return lambda shape: next(it)[:shape[0], :shape[1]]
    got = m.conditional_marginal_log_likelihood(xt, 3, 10, 2, verbose=False, uniform=source())
    exp = om.conditional_marginal_log_likelihood(xt, 3, 10, 2, verbose=False, uniform=source())
    assert np.isfinite(got) and abs(got - exp) <= 2e-2 * abs(exp), (got, exp)


def test_device_gibbs_sampler_known_answer(ctx):
    """The device-resident sampler (pgmvae_model_gibbs_cmll) on the case of test_oracle.test_gibbs_cmll_known_answer,
    small enough to do by hand: dim 3, p1 2 -> blocks [0,1] and [2]; 6 sweep steps, the counter runs for
    i > burn_in * p1 = 2; with p(y=1) = 1 everywhere every draw yields 1 -> counts (1, 2, 3) over denominators (2, 2, 4)
    (core/model.py:132-148, including the float floor division of the last block's denominator)."""
    from core.model import VqVAE
    m = VqVAE([2, 2, 2, 2], 3, 2, 4, seed=1, max_batch=8)
    m.dist = np.ones((3, 4))
    x = np.array([[1, 0, 1]], dtype=np.uint8)
    got = m.conditional_marginal_log_likelihood(x, 2, 3, 1, uniform=lambda sh: np.zeros(sh, np.float32))
    exp = np.log(np.float32(0.5 + 1e-5)) + np.log(np.float32(1 - 1.0 + 1e-5)) + np.log(np.float32(0.75 + 1e-5))
    assert abs(got - exp) < 1e-5, (got, exp)
    # the on-device generator: finite, reproducible for a seed, different for another
    m.dist = np.full((3, 4), 0.5)
    a = m.conditional_marginal_log_likelihood(np.array([[1, 0, 1], [0, 0, 1]], np.uint8), 2, 40, 5, seed=7)
    b = m.conditional_marginal_log_likelihood(np.array([[1, 0, 1], [0, 0, 1]], np.uint8), 2, 40, 5, seed=7)
    c = m.conditional_marginal_log_likelihood(np.array([[1, 0, 1], [0, 0, 1]], np.uint8), 2, 40, 5, seed=8)
    assert np.isfinite(a) and a == b and a != c


def test_save_load_weights_round_trip(ctx, tmp_path):
    """VqVAE.save_weights / load_weights (run.py:63 intent): every tensor in the reference layouts + Adam moments,
    step counter and the EMA debias steps; a model restored from the file continues exactly like the original.
    The path is used as given or with '.npz' appended (np.savez adds the suffix on its own)."""
    from core.model import VqVAE, Adam
    units, V, D, K, B = [15, 14, 13, 12], 16, 4, 32, 128
    y = O.synthetic_binary(4 * B, V, seed=21)
    a = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=True, seed=3, max_batch=B)
    a.compile(optimizer=Adam(lr=1e-3))
    for s in range(2):
        a.train_on_batch(np.ascontiguousarray(y[s * B:(s + 1) * B]))
    path = str(tmp_path / "ckpt")                        # no suffix on purpose
    a.save_weights(path)
    assert (tmp_path / "ckpt.npz").exists()
    b = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=True, seed=99, max_batch=B)
    b.compile(optimizer=Adam(lr=1e-3))
    b.load_weights(path)
    assert b._adam_t == a._adam_t == 2 and b._ema_steps == a._ema_steps == 2
    for n in a.tensor_names():
        np.testing.assert_array_equal(a._get_tensor(n), b._get_tensor(n), err_msg=n)
    for s in range(2, 4):
        ma = a.train_on_batch(np.ascontiguousarray(y[s * B:(s + 1) * B]))
        mb = b.train_on_batch(np.ascontiguousarray(y[s * B:(s + 1) * B]))
        for k in ma:         # (the loss accumulators are fp32 / fp64 atomics: the order of the additions is not fixed)
            assert abs(ma[k] - mb[k]) <= 1e-6 * abs(ma[k]) + 1e-15, (s, k, ma, mb)
    for n in ("fd0.kernel", "fd9.bias", "vq.embeddings", "vq.ema_w", "vq.ema_cluster_size"):
        np.testing.assert_allclose(a._get_tensor(n), b._get_tensor(n), rtol=1e-5, atol=1e-9, err_msg=n)


def test_encode_equals_code_only_one_hot(ctx):
    from core.model import VqVAE
    units, V, D, K, B = [15, 14, 13, 12], 16, 4, 32, 77
    y = O.synthetic_binary(B, V, seed=5)
    m = VqVAE(units, V, D, K, seed=1, max_batch=B)
    idx = m.encode(y)
    oh = m(y, code_only=True)
    assert oh.shape == (V, B, K) and np.array_equal(oh.argmax(-1), idx) and np.all(oh.sum(-1) == 1.0)


def test_tf_crosscheck_export(ctx, tmp_path):
    """tools/tf_crosscheck.py export: the file a TensorFlow install needs to run the unmodified reference on the same
    weights and batches (the check mode itself needs TensorFlow and is not run here)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "case.npz"
    r = subprocess.run([sys.executable, os.path.join(root, "pgm-vae_b200", "tools", "tf_crosscheck.py"), "export", str(out),
                        "--steps", "2", "--eval", "500"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    for k in ("init.fd0.kernel", "init.vq.embeddings", "final.fd9.bias", "final.vq.ema_w", "y_train", "metrics", "eval.idx",
              "eval.n1", "eval.dist", "eval.pll"):
        assert k in z.files, k
    assert z["init.fd0.kernel"].shape == (16, 15, 15) and z["metrics"].shape == (2, 4) and z["eval.idx"].shape == (16, 500)


def test_count_stream_equals_count(ctx, tmp_path):
    """Stage 2 over a stream of chunks (a CSV file read in pieces by a background thread into pinned buffers, and an
    in-memory iterator) equals stage 2 over the whole array: exactly, the counts are integers."""
    from core.model import VqVAE
    from pgmvae import data
    units, V, D, K, B = [15, 14, 13, 12], 16, 4, 32, 256
    y = O.synthetic_binary(5000, V, seed=8)
    path = tmp_path / "toy.train.data"
    with open(path, "w") as f:
        for row in y:
            f.write(",".join(str(int(t)) for t in row) + "\n")
    chunks = list(data.iter_binary_csv(str(path), 700, nvar=V))
    assert sum(len(c) for c in chunks) == len(y) and np.array_equal(np.concatenate(chunks), y)
    m = VqVAE(units, V, D, K, seed=2, max_batch=B)
    n1, n0 = m.count(y)
    for it in (data.iter_binary_csv(str(path), 700, nvar=V), data.iter_array(y, 1234)):
        s1, s0, n = m.count_stream(it, rows_per_chunk=512)
        assert n == len(y) and np.array_equal(s1, n1) and np.array_equal(s0, n0)
    m.dist = (n1 + 0.8) / (n1 + n0 + 1.6)
    assert abs(m.pseudo_log_likelihood_stream(data.iter_array(y, 999), 512) - m.pseudo_log_likelihood(y)) < 1e-12
