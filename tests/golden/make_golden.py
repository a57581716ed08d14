"""Generates tests/golden/*.npz from the CPU oracle (oracle/pgmvae_oracle.py).

The reference ships no golden vectors and TensorFlow is unavailable, so these
fixtures are produced by the oracle itself; they pin the oracle against regressions and give
the GPU parity tests a fixed target.  Run from the repo root:

    python tests/golden/make_golden.py          # needs /root/reference only for the nltcs rows
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "pgm-vae_b200"))
import pgmvae_oracle as O  # noqa: E402

torch.set_num_threads(1)          # deterministic summation order


def make_case(name, units, V, D, K, cost, decay, ema, B, steps, y_all, n_eval, lr, seed):
    params = O.init_params(units, V, D, K, seed=seed)
    m = O.OracleVqVAE(units, V, D, K, cost=cost, decay=decay, ema=ema, params=params)
    out = {"units": np.array(units), "V": V, "D": D, "K": K, "cost": cost, "decay": decay, "ema": int(ema),
           "B": B, "steps": steps, "lr": lr}
    for n, t in params.items():
        out["init." + n] = t.numpy()
    ys = y_all[: B * steps].reshape(steps, B, V)
    out["y_train"] = ys.astype(np.uint8)
    # step 1: activations, codes, gradients (before the update)
    keep = {}
    x0 = O.make_xs(ys[0])
    m2 = O.OracleVqVAE(units, V, D, K, cost=cost, decay=decay, ema=ema, params=params)
    rec = m2(x0, training=False, keep=keep)
    for kname, t in keep.items():
        out["act." + kname] = t.detach().numpy()
    out["act.idx"] = m2.last_idx.numpy().astype(np.int32)
    z = keep["h5"].detach()
    idx, gap = O.vq_assign(z, params["vq.embeddings"])
    out["act.gap"] = gap.numpy()
    out["act.out"] = rec.detach().numpy()
    met, grads = m.loss_and_grads(x0)
    for n, g in grads.items():
        out["grad1." + n] = g.numpy()
    if ema:
        c, dw = O.ema_stats(z, idx, K)
        out["stat1.counts"], out["stat1.dw"] = c.numpy(), dw.numpy()
    # the EMA update of loss_and_grads already ran once on m; rebuild for the real run
    m = O.OracleVqVAE(units, V, D, K, cost=cost, decay=decay, ema=ema, params=params)
    mets = []
    for s in range(steps):
        mets.append(m.train_step(O.make_xs(ys[s]), lr=lr))
        if s in (0, steps - 1):
            tag = f"state{s + 1}."
            for n, a in m.state_numpy().items():
                out[tag + n] = a
            for n in m.trainable:
                out[tag + "adam_m." + n] = m.adam_m[n].numpy().copy()
                out[tag + "adam_v." + n] = m.adam_v[n].numpy().copy()
    out["metrics"] = np.array([[d["loss"], d["mse"], d["mae"], d["vq_loss"]] for d in mets])
    # stage 2 on the trained oracle
    y_eval = y_all[B * steps: B * steps + n_eval]
    x_eval = O.make_xs(y_eval)
    n1, n0 = m.count(x_eval, y_eval)
    m.dist = m.cpt(x_eval, y_eval)
    out["y_eval"] = y_eval.astype(np.uint8)
    out["n1"], out["n0"] = n1.numpy(), n0.numpy()
    out["dist"] = m.dist.numpy()
    out["pll"] = m.pseudo_log_likelihood(x_eval, y_eval)
    m(x_eval[:50], code_only=True)
    out["eval_idx"] = m.last_idx.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: round(v, 6) for k, v in mets[-1].items()}, "pll", out["pll"])


def main():
    nltcs = None
    p = "/root/reference/data/trw/nltcs.train.data"
    if os.path.exists(p):
        nltcs = np.loadtxt(p, delimiter=",", dtype=np.float32).astype(np.uint8)
    else:
        from pgmvae import data
        nltcs = data.load_split("nltcs", "train", 16)
    syn4 = O.synthetic_binary(400, 4, seed=3)
    syn9 = O.synthetic_binary(600, 9, seed=5)
    make_case("v4_ema", [6, 5, 4, 3], 4, 2, 5, 0.25, 0.99, True, B=7, steps=3, y_all=syn4, n_eval=233, lr=1e-2, seed=1)
    make_case("v9_grad", [8, 7, 6, 5], 9, 3, 6, 0.5, 0.99, False, B=33, steps=3, y_all=syn9, n_eval=401, lr=1e-2, seed=2)
    make_case("v16_ema", [15, 14, 13, 12], 16, 4, 32, 0.25, 0.99, True, B=64, steps=3, y_all=nltcs, n_eval=1000,
              lr=1e-3, seed=0)


if __name__ == "__main__":
    main()
