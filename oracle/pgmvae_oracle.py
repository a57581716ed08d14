"""CPU oracle for the pgm-vae hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``pgm-vae_b200/``) never imports it and has no CPU fallback.

What it is: a line-by-line fp32 restatement, on torch-CPU tensors with torch
autograd standing in for TensorFlow's GradientTape, of the reference's

  * packed per-variable dense layer        core/dense.py:99-111
  * VQ layer (gradient-trained)            core/quantizer.py:41-59
  * VQ layer (EMA)                         core/quantizer.py:120-162
  * model wiring, count / cpt / PLL        core/model.py:39-55, 58-96, 99-108
  * leave-one-out input construction       run.py:46-50
  * training configuration                 run.py:59-62  (Keras 'mse' + add_loss, metric 'mae', Adam)

Third-party arithmetic (TensorFlow 2.x, version unpinned by the reference:
README.md:35-37, no lock file; TensorFlow is absent from this image) is
restated from its published definition:

  * Keras MSE / MAE: mean over the last axis, then mean over the rest == global mean;
    total loss = MSE + sum(layer.add_loss terms).
  * Keras Adam (fused ResourceApplyAdam form):
        alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
        m += (g - m) * (1 - b1);  v += (g*g - v) * (1 - b2)
        theta -= alpha * m / (sqrt(v) + eps),     b1=.9  b2=.999  eps=1e-7
  * tf.python.training.moving_averages.assign_moving_average(var, value, decay)
    with its default zero_debias=True: a hidden zero-initialised accumulator
    ``biased`` and a hidden step counter ``local_step`` per EMA'd variable:
        biased -= (biased - value) * (1 - decay);  local_step += 1
        var = biased / (1 - decay^local_step)
    (the same debiased form the author vendored in extern/vqvae.py:72-76,89-95).
  * Keras initialisers on rank-3 shapes [V, a, b]: receptive field = V, so
    fan_in = V*a, fan_out = V*b; he_uniform U(+-sqrt(6/fan_in));
    glorot_uniform U(+-sqrt(6/(fan_in+fan_out)));
    VarianceScaling(scale=1, 'fan_in', 'uniform') U(+-sqrt(3/fan_in)).
  * tf.argmin: lowest index on ties (torch.argmin returns the first minimum too).
  * selu: scale 1.0507009873554805, alpha 1.6732632423543772; sigmoid 1/(1+exp(-x)).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and cannot be executed here (no TensorFlow), so
this oracle is pinned only by (i) internal cross-checks in
tests/test_oracle.py (an independent numpy float64 statement of the same maths,
the leave-one-out identity, count == histogram, PLL == per-sample log-prob mean,
EMA step-1 identity) and (ii) the committed fixtures under tests/golden/ that
it generated itself (tests/golden/make_golden.py).  TF's RNG streams cannot be
matched, so parity is defined on identical injected weights and batch order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

SELU_SCALE = 1.0507009873554805
SELU_ALPHA = 1.6732632423543772
SELU_SCALE_ALPHA = 1.7580993408473768   # scale * alpha, the constant TF precomputes
F32 = torch.float32


# --------------------------------------------------------------------------- #
# run.py:46-50  leave-one-out inputs
# --------------------------------------------------------------------------- #
def make_xs(ys) -> torch.Tensor:
    """run.py:46-50.  ys [N,V] -> xs [N,V,V-1]; row v of a sample is y without element v.

    Restates ``reshape(gather(tile(x,[V]), idx), [V,-1])`` with
    ``idx = [i for i in range(V*V) if i % (V+1) != 0]`` (run.py:46).
    """
    ys = torch.as_tensor(np.asarray(ys), dtype=F32)
    n, v = ys.shape
    idx = torch.tensor([i for i in range(v * v) if i % (v + 1) != 0], dtype=torch.long)
    tiled = ys.repeat(1, v)                      # tf.tile(x, [V]) per row
    return tiled[:, idx].reshape(n, v, v - 1)


# --------------------------------------------------------------------------- #
# initialisers (Keras semantics, our own documented generator)
# --------------------------------------------------------------------------- #
def _uniform(rng: np.random.Generator, shape, limit: float) -> torch.Tensor:
    return torch.from_numpy(rng.uniform(-limit, limit, size=shape).astype(np.float32))


def layer_dims(units: Sequence[int], nvar: int, dim: int) -> List[Tuple[int, int]]:
    """(in, out) of fd0..fd9 (core/model.py:21-36)."""
    u = list(units)
    chain = [nvar - 1, u[0], u[1], u[2], u[3], dim, u[3], u[2], u[1], u[0], nvar - 1]
    return [(chain[i], chain[i + 1]) for i in range(10)]


def init_params(units, nvar, dim, k, seed=0) -> Dict[str, torch.Tensor]:
    """Weights with the reference's initialiser *distributions* (core/model.py:19-36,
    core/quantizer.py:36,112-117), drawn from numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    p: Dict[str, torch.Tensor] = {}
    for i, (fin, fout) in enumerate(layer_dims(units, nvar, dim)):
        fan_in, fan_out = nvar * fin, nvar * fout
        if i < 9:
            lim = math.sqrt(6.0 / fan_in)                 # he_uniform
        else:
            lim = math.sqrt(6.0 / (fan_in + fan_out))     # glorot_uniform
        p[f"fd{i}.kernel"] = _uniform(rng, (nvar, fin, fout), lim)
        p[f"fd{i}.bias"] = torch.zeros(nvar, 1, fout, dtype=F32)
    lim = math.sqrt(3.0 / (nvar * dim))                   # VarianceScaling(1, fan_in, uniform)
    p["vq.embeddings"] = _uniform(rng, (nvar, dim, k), lim)
    return p


# --------------------------------------------------------------------------- #
# core/dense.py:99-111
# --------------------------------------------------------------------------- #
class _TfSelu(torch.autograd.Function):
    """tf.nn.selu as TensorFlow's fused op computes it (Keras activations.selu ->
    nn.selu -> functor::Selu / functor::SeluGrad):
        forward : x < 0 ? scale_alpha * (exp(x) - 1) : scale * x
        backward: out < 0 ? g * (out + scale_alpha) : g * scale     (uses the OUTPUT;
                  at out == 0 the slope is `scale`, which matters for all-zero input rows
                  with zero-initialised biases)."""

    @staticmethod
    def forward(ctx, x):
        out = torch.where(x < 0, np.float32(SELU_SCALE_ALPHA) * (torch.exp(x) - 1.0), np.float32(SELU_SCALE) * x)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        return torch.where(out < 0, g * (out + np.float32(SELU_SCALE_ALPHA)), g * np.float32(SELU_SCALE))


def selu(x: torch.Tensor) -> torch.Tensor:
    return _TfSelu.apply(x)


def fatdense_call(inputs, kernel, bias, activation: Optional[str], fts=None):
    """core/dense.py:99-111: act(inputs[V,B,in] @ kernel[V,in,out] + bias[V,1,out])."""
    if fts is not None:                                   # :104-105
        kernel = kernel.index_select(0, fts)
        bias = bias.index_select(0, fts)
    out = torch.matmul(inputs, kernel)                    # :106
    out = out + bias                                      # :108
    if activation == "selu":                              # :109-110
        return selu(out)
    if activation == "sigmoid":
        return torch.sigmoid(out)
    return out


# --------------------------------------------------------------------------- #
# core/quantizer.py  distances / argmin shared by both VQ layers
# --------------------------------------------------------------------------- #
def vq_distances(inputs, w):
    """core/quantizer.py:44-46 / :135-137, in the reference's association order:
    (sum(z^2) - 2 z@E) + sum(E^2)."""
    return (torch.sum(inputs ** 2, 2, keepdim=True)
            - 2 * torch.matmul(inputs, w)
            + torch.sum(w ** 2, 1, keepdim=True))


def vq_gather(w, idx):
    """tf.gather(tf.transpose(w,[0,2,1]), idx, axis=1, batch_dims=1) -> [V,B,D]."""
    wt = w.transpose(1, 2)                                # [V,K,D]
    return torch.gather(wt, 1, idx.unsqueeze(-1).expand(-1, -1, wt.shape[2]))


class EmaState:
    """Variables of VectorQuantizerEMA (core/quantizer.py:111-117) plus the hidden
    zero-debias accumulators TF creates inside assign_moving_average."""

    def __init__(self, embeddings: torch.Tensor):
        v, d, k = embeddings.shape
        self.ema_cluster_size = torch.zeros(v, k, dtype=F32)       # :113-114
        self.ema_w = embeddings.clone()                            # :116-117
        self.biased_c = torch.zeros(v, k, dtype=F32)               # hidden
        self.biased_w = torch.zeros(v, d, k, dtype=F32)            # hidden
        self.step_c = 0                                            # hidden local_step
        self.step_w = 0


def assign_moving_average(biased: torch.Tensor, step: int, value: torch.Tensor, decay: float):
    """TF moving_averages.assign_moving_average(zero_debias=True), fp32.
    Returns (new_biased, new_step, unbiased_value_assigned_to_variable)."""
    one_minus = torch.tensor(1.0 - decay, dtype=F32)              # convert_to_tensor(1.0 - decay)
    biased = biased - (biased - value) * one_minus
    step = step + 1
    bias_factor = 1 - torch.pow(1.0 - one_minus, torch.tensor(float(step), dtype=F32))
    return biased, step, biased / bias_factor


# --------------------------------------------------------------------------- #
# the model
# --------------------------------------------------------------------------- #
class OracleVqVAE:
    """core/model.py:14-96 on torch-CPU.  Parameters are leaf tensors with
    requires_grad; EMA state and Adam slots are plain tensors."""

    ACTS = ["selu"] * 9 + ["sigmoid"]

    def __init__(self, units, nvar, dim, k, cost=0.5, decay=0.99, ema=True,
                 params: Optional[Dict[str, torch.Tensor]] = None, seed=0, epsilon=1e-5):
        self.units, self.nvar, self.dim, self.k = list(units), nvar, dim, k
        self.cost, self.decay, self.ema, self.epsilon = cost, decay, ema, epsilon
        p = params if params is not None else init_params(units, nvar, dim, k, seed)
        self.p = {n: torch.as_tensor(np.asarray(t), dtype=F32).clone() for n, t in p.items()}
        self.trainable = [f"fd{i}.{s}" for i in range(10) for s in ("kernel", "bias")]
        if not ema:
            self.trainable.append("vq.embeddings")
        for n in self.trainable:
            self.p[n].requires_grad_(True)
        self.ema_state = EmaState(self.p["vq.embeddings"].detach()) if ema else None
        self.losses: List[torch.Tensor] = []
        self.dist = torch.zeros(nvar, k, dtype=torch.float64)     # core/model.py:37
        # Adam slots (run.py:60)
        self.adam_t = 0
        self.adam_m = {n: torch.zeros_like(self.p[n]) for n in self.trainable}
        self.adam_v = {n: torch.zeros_like(self.p[n]) for n in self.trainable}

    # -- layers ------------------------------------------------------------ #
    def _fd(self, i, x, fts):
        return fatdense_call(x, self.p[f"fd{i}.kernel"], self.p[f"fd{i}.bias"], self.ACTS[i], fts)

    def vq_layer(self, inputs, training=None, code_only=False, fts=None):
        emb = self.p["vq.embeddings"]
        w = emb if fts is None else emb.index_select(0, fts)
        distances = vq_distances(inputs, w)
        idx = torch.argmin(distances, 2)
        if self.ema:
            # core/quantizer.py:139-161
            if not code_only:
                quantized = vq_gather(w, idx)
                e_latent = torch.mean((quantized.detach() - inputs) ** 2)
                if training:
                    st = self.ema_state
                    enc = torch.nn.functional.one_hot(idx, self.k).to(F32)
                    with torch.no_grad():
                        st.biased_c, st.step_c, upd_c = assign_moving_average(
                            st.biased_c, st.step_c, enc.sum(1), self.decay)            # :144-145
                        st.ema_cluster_size = upd_c
                        dw = torch.matmul(inputs.detach().transpose(1, 2), enc)          # :146
                        st.biased_w, st.step_w, upd_w = assign_moving_average(
                            st.biased_w, st.step_w, dw, self.decay)                    # :147
                        st.ema_w = upd_w
                        n = upd_c.sum(1, keepdim=True)                                   # :148
                        upd_c = (upd_c + self.epsilon) / (n + self.k * self.epsilon) * n  # :149-150
                        new_w = upd_w / upd_c.unsqueeze(1)                               # :151
                        quantized = quantized.detach()       # gathered before the assign
                        self.p["vq.embeddings"] = new_w                                  # :152
                loss = self.cost * e_latent                                              # :153/155
                output = inputs + (quantized - inputs).detach()                          # :156
            else:
                loss = torch.tensor(0.0)
                output = torch.nn.functional.one_hot(idx, self.k).to(F32) if fts is None else idx
        else:
            # core/quantizer.py:48-56
            if not code_only:
                quantized = vq_gather(w, idx)
                e_latent = torch.mean((quantized.detach() - inputs) ** 2)
                q_latent = torch.mean((quantized - inputs.detach()) ** 2)
                loss = q_latent + self.cost * e_latent
                output = inputs + (quantized - inputs).detach()
            else:
                loss = torch.tensor(0.0)
                output = torch.nn.functional.one_hot(idx, self.k).to(F32) if fts is None else idx
        self.losses.append(loss)
        self.last_idx = idx
        return output

    # -- core/model.py:39-55 ------------------------------------------------ #
    def __call__(self, inputs, training=None, code_only=False, fts=None, keep=None):
        self.losses = []
        x = inputs.transpose(0, 1).contiguous() if fts is None else inputs   # tf.transpose materialises
        for i in range(5):
            x = self._fd(i, x, fts)
            if keep is not None:
                keep[f"h{i + 1}"] = x
        x = self.vq_layer(x, training=training, code_only=code_only, fts=fts)
        if keep is not None:
            keep["vq_out"] = x
        if not code_only:
            for i in range(5, 10):
                x = self._fd(i, x, fts)
                if keep is not None:
                    keep[f"h{i + 1}"] = x
            x = x.transpose(0, 1).contiguous()
        return x

    # -- run.py:60-62  one Keras fit step ----------------------------------- #
    def loss_and_grads(self, x):
        """Forward (training=True) + backward.  Returns (metrics dict, grads dict)."""
        for n in self.trainable:
            self.p[n].grad = None
        out = self(x, training=True)
        mse = torch.mean((out - x) ** 2)
        mae = torch.mean(torch.abs(out - x))
        vq_loss = sum(self.losses)
        loss = mse + vq_loss
        loss.backward()
        grads = {}
        for n in self.trainable:
            g = self.p[n].grad
            grads[n] = torch.zeros_like(self.p[n]) if g is None else g.detach().clone()
        m = {"loss": float(loss.detach()), "mse": float(mse.detach()), "mae": float(mae.detach()),
             "vq_loss": float(vq_loss.detach())}
        return m, grads

    def adam_apply(self, grads, lr, b1=0.9, b2=0.999, eps=1e-7):
        self.adam_t += 1
        t = self.adam_t
        alpha = np.float32(lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
        with torch.no_grad():
            for n in self.trainable:
                if n == "vq.embeddings" and self.ema:
                    continue
                g, m, v = grads[n], self.adam_m[n], self.adam_v[n]
                m += (g - m) * np.float32(1.0 - b1)
                v += (g * g - v) * np.float32(1.0 - b2)
                self.p[n] -= (m * alpha) / (torch.sqrt(v) + np.float32(eps))

    def train_step(self, x, lr=1e-3):
        m, g = self.loss_and_grads(x)
        self.adam_apply(g, lr)
        return m

    def fit(self, x, batch_size, epochs, lr=1e-3, order: Optional[Sequence[np.ndarray]] = None):
        """model.fit(x, x, batch_size, epochs) with an injected per-epoch sample order
        (Keras' shuffle stream is not reproducible without TF).  order[e] is a
        permutation of range(N); None = identity order.  Last partial batch kept."""
        n = x.shape[0]
        hist = []
        for e in range(epochs):
            perm = np.arange(n) if order is None else np.asarray(order[e])
            for s in range(0, n, batch_size):
                hist.append(self.train_step(x[torch.as_tensor(perm[s:s + batch_size])], lr))
        return hist

    # -- core/model.py:58-96 ------------------------------------------------ #
    @torch.no_grad()
    def count(self, x, y, chunk=200):
        """core/model.py:58-82: chunks of 200, one-hot codes masked by y / 1-y."""
        n1 = torch.zeros(1, dtype=F32)
        n0 = torch.zeros(1, dtype=F32)
        y = torch.as_tensor(np.asarray(y), dtype=F32)
        for s in range(0, y.shape[0], chunk):
            y_ = y[s:s + chunk].t()                                   # [V,b]
            code = self(x[s:s + chunk], code_only=True)               # [V,b,K]
            n1 = n1 + (code * (y_ != 0).to(F32).unsqueeze(-1)).sum(1)  # boolean_mask + reduce_sum
            n0 = n0 + (code * ((1 - y_) != 0).to(F32).unsqueeze(-1)).sum(1)
        return n1.to(torch.float64), n0.to(torch.float64)

    def cpt(self, x, y):
        n1, n0 = self.count(x, y)
        return (n1 + 0.8) / (n1 + n0 + 1.6)                            # core/model.py:88

    def pseudo_log_likelihood(self, x, y):
        lp1 = torch.log(self.dist + 1e-5)                              # :93
        lp0 = torch.log(1 - self.dist + 1e-5)                          # :94
        n1, n0 = self.count(x, y)
        return float(torch.sum(n1 * lp1 + n0 * lp0) / y.shape[0])      # :96

    @torch.no_grad()
    def get_probability(self, x, fts):
        """core/model.py:99-108."""
        fts = torch.as_tensor(fts, dtype=torch.long)
        enc_idx = self(x, code_only=True, fts=fts)                     # [F,B]
        prb = self.dist.index_select(0, fts).to(F32)
        return torch.gather(prb, 1, enc_idx)

    def conditional_marginal_log_likelihood(self, x, p1, num_smp, burn_in, verbose=False, uniform=None):
        """core/model.py:110-148."""
        rng = np.random.default_rng(0)
        uni = uniform if uniform is not None else (lambda shape: rng.random(shape, dtype=np.float32))
        gp = lambda xs, fts: self.get_probability(torch.from_numpy(np.ascontiguousarray(xs)), fts).numpy()
        return gibbs_cmll(gp, x, p1, num_smp, burn_in, uni, verbose)

    # -- helpers for tests -------------------------------------------------- #
    def state_numpy(self) -> Dict[str, np.ndarray]:
        out = {n: t.detach().numpy().copy() for n, t in self.p.items()}
        if self.ema:
            st = self.ema_state
            out.update({"vq.ema_cluster_size": st.ema_cluster_size.numpy().copy(),
                        "vq.ema_w": st.ema_w.numpy().copy(),
                        "vq.biased_c": st.biased_c.numpy().copy(),
                        "vq.biased_w": st.biased_w.numpy().copy()})
        return out


# --------------------------------------------------------------------------- #
# Gibbs-sampling conditional marginal log-likelihood (core/model.py:110-148)
# --------------------------------------------------------------------------- #
def gibbs_cmll(get_probability, x, p1, num_smp, burn_in, uniform, verbose=False):
    """Block Gibbs sampler of the reference (core/model.py:110-148), framework-free.

    get_probability(xs [F,B,V-1], fts [F]) -> p(y=1) [F,B];  uniform(shape) -> U[0,1) draws (the reference uses
    tf.random.uniform, whose stream cannot be reproduced without TensorFlow; parity runs inject the draws).
    Keeps the reference's quirks: the counter starts at i > burn_in * p1 (strict), every block sweeps its own
    variables with period vol[b], and the last block's denominator is valid * p1 // vol[-1]."""
    x = np.asarray(x, dtype=np.float32)
    batch_size, dim = x.shape
    blocks = int(np.ceil(dim / p1))                                        # :123
    vol = np.array([p1] * (blocks - 1) + [dim - p1 * (blocks - 1)])        # :124
    marker = np.arange(blocks) * p1                                        # :126
    state = np.tile(x[None], (blocks, 1, 1))                               # :127
    cnt = np.zeros_like(x)                                                 # :128
    for i in range(num_smp * p1):                                          # :132
        y = marker + np.mod(i, vol)                                        # :133
        xs = np.stack([np.delete(state[b], y[b], axis=1) for b in range(blocks)])   # :134-136
        prb = np.asarray(get_probability(xs, y), dtype=np.float32)         # :137
        gibbs = (np.asarray(uniform((blocks, batch_size)), dtype=np.float32) < prb).astype(np.float32)   # :138
        for b in range(blocks):
            state[b, :, y[b]] = gibbs[b]                                   # :139
            if i > burn_in * p1:
                cnt[:, y[b]] += gibbs[b]                                   # :140-141
        if verbose:
            print(f"# of samples: {i // p1}, component: {y[0]}")
    valid = num_smp - burn_in                                              # :146
    valid_end = np.float32(valid * p1) // np.float32(vol[-1])              # :147
    den = np.concatenate([np.full(dim - vol[-1], valid, np.float32), np.full(vol[-1], valid_end, np.float32)])
    cmll = cnt / den[None, :]                                              # :148
    return float(np.sum(x * np.log(cmll + 1e-5) + (1 - x) * np.log(1 - cmll + 1e-5)) / batch_size)   # :149


# --------------------------------------------------------------------------- #
# standalone operator statements used by operator-level parity tests
# --------------------------------------------------------------------------- #
@torch.no_grad()
def vq_assign(z, emb):
    """indices + (best, second-best) distances; z [V,B,D], emb [V,D,K]."""
    d = vq_distances(torch.as_tensor(z, dtype=F32), torch.as_tensor(emb, dtype=F32))
    idx = torch.argmin(d, 2)
    top2 = torch.topk(d, 2 if d.shape[2] > 1 else 1, dim=2, largest=False).values
    gap = (top2[..., -1] - top2[..., 0]) if d.shape[2] > 1 else torch.full(idx.shape, float("inf"))
    return idx, gap


@torch.no_grad()
def ema_stats(z, idx, k):
    """core/quantizer.py:144-146: counts [V,K], dw [V,D,K] via one-hot matmul."""
    z = torch.as_tensor(z, dtype=F32)
    enc = torch.nn.functional.one_hot(torch.as_tensor(idx, dtype=torch.long), k).to(F32)
    return enc.sum(1), torch.matmul(z.transpose(1, 2), enc)


def pll_from_counts(n1, n0, dist, n):
    """core/model.py:93-96 in float64."""
    n1 = np.asarray(n1, dtype=np.float64)
    n0 = np.asarray(n0, dtype=np.float64)
    dist = np.asarray(dist, dtype=np.float64)
    return float(np.sum(n1 * np.log(dist + 1e-5) + n0 * np.log(1 - dist + 1e-5)) / n)


def synthetic_binary(n, v, seed=0) -> np.ndarray:
    """column-wise Bernoulli, densities U(0.02,0.5) from default_rng(seed+1)."""
    dens = np.random.default_rng(seed + 1).uniform(0.02, 0.5, size=v)
    return (np.random.default_rng(seed).random((n, v)) < dens).astype(np.uint8)
