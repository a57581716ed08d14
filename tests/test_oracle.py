"""CPU tests of the oracle (no GPU): regression against the committed fixtures, and
cross-checks against an independent float64 statement of the same maths."""
import glob
import os

import numpy as np
import pytest
import torch

import pgmvae_oracle as O
import np_expanded as NE

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = dict(units=[int(u) for u in z["units"]], V=int(z["V"]), D=int(z["D"]), K=int(z["K"]), cost=float(z["cost"]),
               decay=float(z["decay"]), ema=bool(int(z["ema"])), B=int(z["B"]), steps=int(z["steps"]), lr=float(z["lr"]))
    params = {k[5:]: z[k] for k in z.files if k.startswith("init.")}
    return z, cfg, params


def make_oracle(cfg, params):
    return O.OracleVqVAE(cfg["units"], cfg["V"], cfg["D"], cfg["K"], cost=cfg["cost"], decay=cfg["decay"],
                         ema=cfg["ema"], params={k: torch.from_numpy(v) for k, v in params.items()})


def test_golden_present():
    assert {"v4_ema", "v9_grad", "v16_ema"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    z, cfg, params = load_case(name)
    m = make_oracle(cfg, params)
    mets = []
    for s in range(cfg["steps"]):
        d = m.train_step(O.make_xs(z["y_train"][s]), lr=cfg["lr"])
        mets.append([d["loss"], d["mse"], d["mae"], d["vq_loss"]])
        if s == 0:
            for n, a in m.state_numpy().items():
                np.testing.assert_allclose(a, z["state1." + n], rtol=2e-5, atol=1e-7, err_msg=n)
    np.testing.assert_allclose(np.array(mets), z["metrics"], rtol=1e-5)
    for n, a in m.state_numpy().items():
        np.testing.assert_allclose(a, z[f"state{cfg['steps']}." + n], rtol=5e-5, atol=1e-7, err_msg=n)
    y_eval = z["y_eval"]
    x_eval = O.make_xs(y_eval)
    n1, n0 = m.count(x_eval, y_eval)
    np.testing.assert_array_equal(n1.numpy(), z["n1"])
    np.testing.assert_array_equal(n0.numpy(), z["n0"])
    m.dist = m.cpt(x_eval, y_eval)
    assert abs(m.pseudo_log_likelihood(x_eval, y_eval) - float(z["pll"])) < 1e-9


def test_make_xs_is_leave_one_out():
    # run.py:46-50: row v of a sample is y without element v
    y = np.arange(10, 15, dtype=np.float32)[None].repeat(3, 0) + np.arange(3, dtype=np.float32)[:, None] * 100
    xs = O.make_xs(y).numpy()
    assert xs.shape == (3, 5, 4)
    for n in range(3):
        for v in range(5):
            np.testing.assert_array_equal(xs[n, v], np.delete(y[n], v))
    np.testing.assert_array_equal(xs[0], [[11, 12, 13, 14], [10, 12, 13, 14], [10, 11, 13, 14], [10, 11, 12, 14],
                                          [10, 11, 12, 13]])


@pytest.mark.parametrize("name", CASES)
def test_expanded_formulation_matches_autograd(name):
    """The masked/expanded GEMM formulation with the hand-written backward (what the CUDA
    kernels implement) equals the oracle's autograd on the materialised inputs."""
    z, cfg, params = load_case(name)
    m = make_oracle(cfg, params)
    y = z["y_train"][0]
    met_o, g_o = m.loss_and_grads(O.make_xs(y))
    met_n, g_n, aux = NE.step_grads(params, y, cfg["D"], cfg["K"], cfg["cost"], cfg["ema"])
    for k in ("loss", "mse", "mae", "vq_loss"):
        assert abs(met_o[k] - met_n[k]) <= 2e-6 * max(1.0, abs(met_n[k])), k
    np.testing.assert_array_equal(aux["idx"], z["act.idx"])
    for n, g in g_o.items():
        ref = g_n[n]
        err = np.abs(g.numpy() - ref).max()
        assert err <= 2e-5 * max(np.abs(ref).max(), 1e-8) + 1e-9, (n, err, np.abs(ref).max())
        np.testing.assert_allclose(g.numpy(), z["grad1." + n], rtol=1e-4, atol=1e-9)


def test_selu_follows_tf_at_zero():
    # all-zero rows with zero biases give pre-activations of exactly 0: TF's SeluGrad uses `scale` there
    x = torch.zeros(3, requires_grad=True)
    O.selu(x).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), O.SELU_SCALE, rtol=1e-7)
    x = torch.tensor([-1.0, 2.0], requires_grad=True)
    out = O.selu(x)
    out.sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), [O.SELU_SCALE_ALPHA * (np.exp(-1) - 1), 2 * O.SELU_SCALE], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), [O.SELU_SCALE_ALPHA * np.exp(-1), O.SELU_SCALE], rtol=1e-6)


def test_count_is_histogram_and_pll_is_mean_logprob():
    z, cfg, params = load_case("v4_ema")
    m = make_oracle(cfg, params)
    y = z["y_eval"]
    x = O.make_xs(y)
    n1, n0 = m.count(x, y)
    m(x, code_only=True)
    idx = m.last_idx.numpy()                                   # [V,N]
    h1 = np.zeros((cfg["V"], cfg["K"]))
    h0 = np.zeros_like(h1)
    for v in range(cfg["V"]):
        np.add.at(h1[v], idx[v][y[:, v] != 0], 1)
        np.add.at(h0[v], idx[v][y[:, v] == 0], 1)
    np.testing.assert_array_equal(n1.numpy(), h1)
    np.testing.assert_array_equal(n0.numpy(), h0)
    assert (h1 + h0).sum() == y.shape[0] * cfg["V"]
    m.dist = m.cpt(x, y)
    pll = m.pseudo_log_likelihood(x, y)
    d = m.dist.numpy()
    p = d[np.arange(cfg["V"])[:, None], idx]                  # p(y_v=1 | code)
    lp = np.where(y.T != 0, np.log(p + 1e-5), np.log(1 - p + 1e-5))
    assert abs(pll - lp.sum(0).mean()) < 1e-9
    assert abs(pll - O.pll_from_counts(h1, h0, d, y.shape[0])) < 1e-12


def test_ema_first_step_identity():
    """zero-debiased EMA: after the first update the visible averages equal the raw statistics."""
    z, cfg, params = load_case("v4_ema")
    m = make_oracle(cfg, params)
    x = O.make_xs(z["y_train"][0])
    keep = {}
    m(x, training=True, keep=keep)
    c, dw = O.ema_stats(keep["h5"].detach(), m.last_idx, cfg["K"])
    np.testing.assert_allclose(m.ema_state.ema_cluster_size.numpy(), c.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.ema_state.ema_w.numpy(), dw.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(c.numpy().sum(1), x.shape[0])
    # unused codes collapse to (almost) zero vectors, used ones to the Laplace-normalised mean
    E = m.p["vq.embeddings"].numpy()
    used = c.numpy() > 0
    n, K, eps = c.numpy().sum(1, keepdims=True), cfg["K"], 1e-5
    size = (c.numpy() + eps) / (n + K * eps) * n
    np.testing.assert_allclose(E, dw.numpy() / size[:, None, :], rtol=1e-4, atol=1e-6)
    assert np.all(E.transpose(0, 2, 1)[~used] == 0)


def test_argmin_lowest_index_on_ties():
    zv = torch.zeros(1, 3, 2)
    emb = torch.tensor([[[1.0, 1.0, 0.5, 1.0], [0.0, 0.0, 0.5, 0.0]]])     # codes 0,1,3 identical
    idx, gap = O.vq_assign(zv, emb)
    assert idx.tolist() == [[2, 2, 2]]
    emb2 = torch.tensor([[[1.0, 1.0, 2.0, 1.0], [0.0, 0.0, 0.5, 0.0]]])
    idx, gap = O.vq_assign(zv, emb2)
    assert idx.tolist() == [[0, 0, 0]] and float(gap.max()) == 0.0


def test_adam_matches_torch_formula():
    rng = np.random.default_rng(0)
    params = O.init_params([3, 3, 3, 3], 3, 2, 4, seed=0)
    m = O.OracleVqVAE([3, 3, 3, 3], 3, 2, 4, ema=True, params=params)
    y = O.synthetic_binary(8, 3, seed=0)
    p0 = m.p["fd2.kernel"].detach().clone().double()
    met, g = m.loss_and_grads(O.make_xs(y))
    m.adam_apply(g, 1e-2)
    gg = g["fd2.kernel"].double()
    mm, vv = 0.1 * gg, 0.001 * gg * gg
    alpha = 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = p0 - alpha * mm / (vv.sqrt() + 1e-7)
    np.testing.assert_allclose(m.p["fd2.kernel"].detach().numpy(), exp.numpy(), rtol=1e-5, atol=1e-7)


def test_gibbs_cmll_known_answer():
    """core/model.py:110-148 on a case small enough to do by hand: dim 3, p1 2 -> blocks [0,1] and [2];
    6 sweep steps, the counter runs for i > burn_in * p1 = 2; every draw yields 1 -> counts (1, 2, 3) over
    denominators (2, 2, 4)."""
    x = np.array([[1, 0, 1]], dtype=np.float32)
    got = O.gibbs_cmll(lambda xs, fts: np.ones(xs.shape[:2], np.float32), x, 2, 3, 1, lambda sh: np.zeros(sh, np.float32))
    exp = np.log(0.5 + 1e-5) + np.log(1 - 1.0 + 1e-5) + np.log(0.75 + 1e-5)
    assert abs(got - exp) < 1e-5


def test_tf_crosscheck_make_xs_matches_run_py():
    """tools/tf_crosscheck.py builds the reference's leave-one-out inputs in numpy; same tensor as run.py:46-50."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tf_crosscheck", os.path.join(root, "pgm-vae_b200", "tools", "tf_crosscheck.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    y = O.synthetic_binary(9, 7, seed=3)
    np.testing.assert_array_equal(mod.make_xs_np(y), O.make_xs(y).numpy())
