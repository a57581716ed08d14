"""Independent float64 numpy statement of the EXPANDED formulation the CUDA path uses
(test infrastructure): net v reads the raw matrix y[B,V] through a weight with row v zeroed,
reconstructs all V columns with column v masked, and the backward pass is written out by
hand exactly as the kernels fuse it.  Checked against the oracle's autograd in
tests/test_oracle.py so that the kernel design is validated on the CPU first."""
import numpy as np

SCALE = 1.0507009873554804934193349852946
SCALE_ALPHA = 1.7580993408473768599402175208123


def selu(x):
    return np.where(x < 0, SCALE_ALPHA * (np.exp(np.minimum(x, 0)) - 1.0), SCALE * x)


def dselu_from_out(h):
    return np.where(h < 0, h + SCALE_ALPHA, SCALE)


def expand_params(p, V):
    """reference-layout params -> expanded float64 params (W0 [V,V,u0] zero row v; W9 [V,u0,V], b9 [V,V])."""
    q = {k: np.asarray(v, dtype=np.float64) for k, v in p.items()}
    W0 = np.zeros((V, V, q["fd0.kernel"].shape[2]))
    W9 = np.zeros((V, q["fd9.kernel"].shape[1], V))
    b9 = np.zeros((V, 1, V))
    for v in range(V):
        keep = [j for j in range(V) if j != v]
        W0[v, keep, :] = q["fd0.kernel"][v]
        W9[v][:, keep] = q["fd9.kernel"][v]
        b9[v, 0, keep] = q["fd9.bias"][v, 0]
    q["fd0.kernel"], q["fd9.kernel"], q["fd9.bias"] = W0, W9, b9
    return q


def contract_grads(g, V):
    out = dict(g)
    out["fd0.kernel"] = np.stack([np.delete(g["fd0.kernel"][v], v, axis=0) for v in range(V)])
    out["fd9.kernel"] = np.stack([np.delete(g["fd9.kernel"][v], v, axis=1) for v in range(V)])
    out["fd9.bias"] = np.stack([np.delete(g["fd9.bias"][v], v, axis=1) for v in range(V)])
    return out


def step_grads(p_ref, y, D, K, cost, ema, global_B=None):
    """Returns (metrics, grads in reference layout, aux) for one batch y [B,V] in {0,1}."""
    y = np.asarray(y, dtype=np.float64)
    B, V = y.shape
    gB = global_B or B
    p = expand_params(p_ref, V)
    H = []
    x = np.broadcast_to(y, (V, B, V))
    for l in range(5):
        x = selu(x @ p[f"fd{l}.kernel"] + p[f"fd{l}.bias"])
        H.append(x)
    z = H[4]
    E = p["vq.embeddings"]                                            # [V,D,K]
    dist = (np.sum(z ** 2, 2, keepdims=True) - 2 * (z @ E)) + np.sum(E ** 2, 1, keepdims=True)
    idx = np.argmin(dist, 2)
    q = np.take_along_axis(E.transpose(0, 2, 1), idx[..., None], axis=1)
    n_lat, n_out = gB * V * D, gB * V * (V - 1)
    e_latent = np.sum((q - z) ** 2) / n_lat
    st = z + (q - z)
    x = st
    for l in range(5, 9):
        x = selu(x @ p[f"fd{l}.kernel"] + p[f"fd{l}.bias"])
        H.append(x)
    o = 1.0 / (1.0 + np.exp(-(H[8] @ p["fd9.kernel"] + p["fd9.bias"])))
    mask = 1.0 - np.eye(V)[:, None, :]                                # [V,1,V] column v masked
    diff = (o - y[None]) * mask
    mse, mae = np.sum(diff ** 2) / n_out, np.sum(np.abs(diff)) / n_out
    vq = cost * e_latent if ema else (1 + cost) * e_latent
    g = {}
    d = (2.0 / n_out) * diff * o * (1 - o)                            # dpre9
    for l in range(9, -1, -1):
        xin = np.broadcast_to(y, (V, B, V)) if l == 0 else (st if l == 5 else H[l - 1])
        g[f"fd{l}.kernel"] = xin.transpose(0, 2, 1) @ d
        g[f"fd{l}.bias"] = d.sum(1, keepdims=True)
        if l > 0:
            dx = d @ p[f"fd{l}.kernel"].transpose(0, 2, 1)
            if l == 5:
                dx = dx + (cost * 2.0 / n_lat) * (z - q)
            d = dx * dselu_from_out(H[l - 1])
    for v in range(V):
        g["fd0.kernel"][v, v, :] = 0.0
    if not ema:
        onehot = np.eye(K)[idx]                                        # [V,B,K]
        g["vq.embeddings"] = ((2.0 / n_lat) * (q - z)).transpose(0, 2, 1) @ onehot
    aux = {"idx": idx, "z": z, "q": q, "out": o, "H": H}
    return {"loss": mse + vq, "mse": mse, "mae": mae, "vq_loss": vq}, contract_grads(g, V), aux
