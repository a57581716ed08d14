"""Cross-check of the B200 implementation against the UNMODIFIED TensorFlow reference (motionlife/pgm-vae).

The reference ships no tests or golden vectors and needs TensorFlow 2.x, which cannot be installed where this
repository is built (no network, no wheel): the oracle in oracle/ is therefore "parity unpinned" (DESIGN.md section 2).
This tool closes that gap for anyone who has TensorFlow: it moves a complete, self-describing test case through a
file.

  1. on a B200 box (this repository):
         python pgm-vae_b200/tools/tf_crosscheck.py export case.npz [--precision fp32] [--nvar 16 --units 15,14,13,12 ...]
     initial weights in the reference's layouts ([V,in,units], [V,1,units], [V,D,K]), the batches (as y [B,V]), and what
     this implementation computed from them: per-step loss / mse / mae / vq_loss of `steps` optimiser steps, the
     weights and EMA state afterwards, the codes, n1 / n0, the CPT and the PLL of an evaluation set.

  2. anywhere with TensorFlow 2.x and a checkout of the reference:
         python tf_crosscheck.py check case.npz --reference /path/to/pgm-vae
     builds the reference's own core.model.VqVAE (nothing of this repository is imported), injects the initial weights,
     feeds the same batches in the same order through model.train_on_batch(x, x) (the Keras fit step of run.py:60-62 with
     the leave-one-out inputs of run.py:46-50), then model.count / cpt / pseudo_log_likelihood, and compares with the
     exported numbers: losses, weights, EMA codebook and PLL within 1e-3 relative, codes bit-exact where the top-2
     distance gap exceeds 1e-5 (north_star's bars).

The check mode has not been run by the authors of this repository (no TensorFlow here); it only uses the reference's
public surface: VqVAE(units, nvar, dim, k, cost, decay, ema), compile/train_on_batch, layer.kernel / .bias,
vq_layer.embeddings / ema_w / ema_cluster_size, count, cpt, pseudo_log_likelihood.
"""
import argparse
import os
import sys

import numpy as np


def make_xs_np(y):
    """run.py:46-50 in numpy: xs[n, v, :] = y[n] without element v."""
    n, v = y.shape
    keep = ~np.eye(v, dtype=bool)
    return np.broadcast_to(y[:, None, :], (n, v, v))[:, keep].reshape(n, v, v - 1).astype(np.float32)


def export(args):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    from pgmvae import _ffi, data
    from core.model import VqVAE, Adam
    units = [int(u) for u in args.units.split(",")]
    V, D, K, B = args.nvar, args.dim, args.embedding, args.batch
    ctx = _ffi.get_context(0)
    ctx.set_precision({"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32, "bf16": _ffi.PREC_BF16}[args.precision])
    m = VqVAE(units, V, D, K, cost=args.cost, decay=args.decay, ema=not args.no_ema, seed=args.seed, max_batch=B)
    m.compile(optimizer=Adam(lr=args.rate))
    out = {"cfg.units": np.array(units), "cfg.nvar": V, "cfg.dim": D, "cfg.k": K, "cfg.batch": B, "cfg.steps": args.steps,
           "cfg.rate": args.rate, "cfg.cost": args.cost, "cfg.decay": args.decay, "cfg.ema": int(not args.no_ema),
           "cfg.precision": args.precision}
    for n, a in m.state_dict().items():
        if not n.startswith(("adam_", "_")):
            out["init." + n] = a
    y = data.synthetic_binary(B * args.steps, V, seed=args.seed + 1)
    out["y_train"] = y.reshape(args.steps, B, V)
    mets = []
    for s in range(args.steps):
        d = m.train_on_batch(np.ascontiguousarray(y[s * B:(s + 1) * B]))
        mets.append([d["loss"], d["mse"], d["mae"], d["vq_loss"]])
    out["metrics"] = np.array(mets)
    for n, a in m.state_dict().items():
        if not n.startswith(("adam_", "_")):
            out["final." + n] = a
    ye = data.synthetic_binary(args.eval, V, seed=args.seed + 2)
    out["y_eval"] = ye
    out["eval.idx"] = m(ye, code_only=True).argmax(-1).astype(np.int32)              # [V, N]
    n1, n0 = m.count(ye)
    m.dist = (n1 + 0.8) / (n1 + n0 + 1.6)
    out["eval.n1"], out["eval.n0"], out["eval.dist"] = n1, n0, m.dist
    out["eval.pll"] = m.pseudo_log_likelihood(ye)
    np.savez(args.file, **out)
    print(f"wrote {args.file}: {args.steps} steps of batch {B}, V={V} units={units} D={D} K={K}, losses {mets[-1]}, "
          f"pll {out['eval.pll']:.6f}")


def check(args):
    sys.path.insert(0, args.reference)
    import tensorflow as tf
    from core.model import VqVAE                                   # the reference's own model (core/model.py:14)
    z = np.load(args.file)
    units, V, D, K = [int(u) for u in z["cfg.units"]], int(z["cfg.nvar"]), int(z["cfg.dim"]), int(z["cfg.k"])
    ema = bool(int(z["cfg.ema"]))
    model = VqVAE(units=units, nvar=V, dim=D, k=K, cost=float(z["cfg.cost"]), decay=float(z["cfg.decay"]), ema=ema)
    model.compile(optimizer=tf.keras.optimizers.Adam(learning_rate=float(z["cfg.rate"])), loss="mse", metrics=["mae"])
    ys = z["y_train"].astype(np.float32)
    model(tf.constant(make_xs_np(ys[0][:2])), training=False)      # builds every layer
    layers = [getattr(model, f"fd{i}") for i in range(10)]
    for i, l in enumerate(layers):
        l.kernel.assign(z[f"init.fd{i}.kernel"])
        l.bias.assign(z[f"init.fd{i}.bias"])
    model.vq_layer.embeddings.assign(z["init.vq.embeddings"])
    if ema:
        model.vq_layer.ema_w.assign(z["init.vq.embeddings"])       # core/quantizer.py:117
    worst = {}

    def rel(name, got, exp):
        got, exp = np.asarray(got, np.float64), np.asarray(exp, np.float64)
        e = float(np.abs(got - exp).max() / max(np.abs(exp).max(), 1e-30))
        worst[name] = max(worst.get(name, 0.0), e)
        return e
    for s in range(int(z["cfg.steps"])):
        x = tf.constant(make_xs_np(ys[s]))
        res = model.train_on_batch(x, x, return_dict=True)         # one Keras fit step (run.py:62)
        rel("loss", res["loss"], z["metrics"][s][0])
        rel("mae", res["mae"], z["metrics"][s][2])
        print(f"step {s}: reference loss {res['loss']:.8f} mae {res['mae']:.8f} | exported {z['metrics'][s][0]:.8f} "
              f"{z['metrics'][s][2]:.8f}")
    for i, l in enumerate(layers):
        rel("weights", l.kernel.numpy(), z[f"final.fd{i}.kernel"])
        rel("weights", l.bias.numpy(), z[f"final.fd{i}.bias"])
    rel("codebook", model.vq_layer.embeddings.numpy(), z["final.vq.embeddings"])
    ye = z["y_eval"].astype(np.float32)
    xe = tf.constant(make_xs_np(ye))
    code = model(xe, code_only=True).numpy()                       # one-hot [V, N, K]
    idx = code.argmax(-1)
    mism = float((idx != z["eval.idx"]).mean())
    n1, n0 = model.count(xe, tf.constant(ye))
    rel("n1", n1.numpy(), z["eval.n1"])
    model.dist = model.cpt(xe, tf.constant(ye))
    rel("dist", model.dist.numpy(), z["eval.dist"])
    pll = float(model.pseudo_log_likelihood(xe, tf.constant(ye)))
    rel("pll", pll, float(z["eval.pll"]))
    print(f"code mismatches {mism:.3e} (must be 0 outside the 1e-5 distance-gap band), pll {pll:.8f} vs {float(z['eval.pll']):.8f}")
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    print("deviations (max-norm relative):", {k: f"{v:.2e}" for k, v in worst.items()})
    print("PARITY OK (1e-3)" if not bad and mism < 1e-3 else f"PARITY FAILED: {bad} code mismatches {mism}")
    return 0 if not bad and mism < 1e-3 else 1


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    e = sub.add_parser("export")
    e.add_argument("file")
    e.add_argument("--nvar", type=int, default=16)
    e.add_argument("--units", default="15,14,13,12")
    e.add_argument("--dim", type=int, default=4)
    e.add_argument("--embedding", type=int, default=32)
    e.add_argument("--batch", type=int, default=256)
    e.add_argument("--steps", type=int, default=3)
    e.add_argument("--eval", type=int, default=2000)
    e.add_argument("--rate", type=float, default=1e-3)
    e.add_argument("--cost", type=float, default=0.25)
    e.add_argument("--decay", type=float, default=0.99)
    e.add_argument("--no-ema", action="store_true")
    e.add_argument("--seed", type=int, default=0)
    e.add_argument("--precision", default="fp32", choices=["fp32", "tf32", "bf16"])
    c = sub.add_parser("check")
    c.add_argument("file")
    c.add_argument("--reference", required=True, help="checkout of motionlife/pgm-vae (unmodified)")
    args = ap.parse_args()
    sys.exit(export(args) if args.cmd == "export" else check(args))


if __name__ == "__main__":
    main()
