"""Data-parallel parity check, run under torchrun on N GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        pgm-vae_b200/tools/dp_check.py
Every rank trains on its share of each global batch through the data-parallel path of libpgmvae.so (NCCL all-reduce
of gradients, EMA statistics and loss accumulators, overlapped with compute; peer-to-peer exchange fused with Adam
at two ranks; per-group sharded peer-to-peer exchange fused with Adam for wide models); rank 0 then replays the same global batches on ONE GPU and the two runs must agree: losses, weights,
codebook, PLL counts (``dp_parity``).  ``main`` walks every schedule the library has: precision fp32 / tf32 (chain
kernels, comm-stream overlap) / bf16 (layer-by-layer tensor-core path, several variable groups),
PGMVAE_P2P in {0, 1}, PGMVAE_DP_BUCKETS in {0, 1}.  bench.py calls ``dp_parity`` on the benchmarked configuration
outside its timed region, so every multi-GPU bench line carries this evidence."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgmvae import _ffi, data, dist  # noqa: E402
from core.model import VqVAE, Adam  # noqa: E402

PREC = {"fp32": _ffi.PREC_FP32, "tf32": _ffi.PREC_TF32, "bf16": _ffi.PREC_BF16}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dp_parity(ctx, comm, rank, world, units, V, D, K, per_rank_batch, precision, ema=True, steps=3, seed=5, local=0,
              tensors=("fd1.kernel", "fd4.bias", "fd9.bias"), eval_samples=None):
    """Train `steps` steps data-parallel (every rank) and on one GPU (rank 0, same global batches, same seed) and
    return the deviations {loss_rel, weight_rel, codebook_rel, count_l1, ...} (None on ranks != 0)."""
    import torch.distributed as tdist
    GB = per_rank_batch * world
    y = data.synthetic_binary(GB * steps, V, seed=seed + 6)
    old = ctx.get_precision() if hasattr(ctx, "get_precision") else None
    ctx.set_precision(PREC[precision])
    try:
        m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, seed=seed, max_batch=per_rank_batch + 1, device=local, comm=comm)
        m.compile(optimizer=Adam(lr=1e-3))
        losses = []
        for s in range(steps):
            gb = y[s * GB:(s + 1) * GB]
            lo, hi = dist.shard_bounds(GB, rank, world)
            met = m.train_on_batch(np.ascontiguousarray(gb[lo:hi]), global_batch=GB)
            losses.append([met["loss"], met["mse"], met["mae"], met["vq_loss"]])
        ye = y if eval_samples is None else y[:eval_samples]
        lo, hi = dist.shard_bounds(len(ye), rank, world)
        n1, n0 = m.count(ye[lo:hi])                       # sample-sharded, counts summed over ranks
        m.dist = (n1 + 0.8) / (n1 + n0 + 1.6)
        pll_s = m.pseudo_log_likelihood(ye[lo:hi], total=len(ye))
        pll_v = m.pseudo_log_likelihood(ye, shard="variables")      # variable-sharded: one scalar crosses ranks
        # (sharded peer-to-peer exchange: the Adam moments and fp32 master kernels live with the owner of a shard until this
        # collective gather)
        lib = _ffi.lib()
        sharded = bool(lib.pgmvae_model_p2p_state_sharded(m._h))
        if sharded:
            _ffi.check(lib.pgmvae_model_p2p_sync_state(m._h))
        moments = ("adam_m.fd1.kernel", "adam_v.fd9.bias")
        mine = {n: m._get_tensor(n) for n in tuple(tensors) + moments}
        got = mine if rank == 0 else None
        emb = m._get_tensor("vq.embeddings") if rank == 0 else None
        # the replicas must be bit-identical: compare a digest of every rank's tensors
        import hashlib
        digest = hashlib.sha1(b"".join(np.ascontiguousarray(mine[n]).tobytes() for n in sorted(mine))).hexdigest()
        digests = [None] * world
        if world > 1:
            tdist.all_gather_object(digests, digest)
        else:
            digests = [digest]
        arith = int(_ffi.lib().pgmvae_model_arithmetic(m._h))
        groups = -(-V // m.group_size())
        del m
        res = None
        if rank == 0:
            ref = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, seed=seed, max_batch=GB, device=local, comm=None)
            ref.compile(optimizer=Adam(lr=1e-3))
            rl = []
            for s in range(steps):
                met = ref.train_on_batch(np.ascontiguousarray(y[s * GB:(s + 1) * GB]))
                rl.append([met["loss"], met["mse"], met["mae"], met["vq_loss"]])
            r1, r0 = ref.count(ye)
            ref.dist = (r1 + 0.8) / (r1 + r0 + 1.6)
            rpll = ref.pseudo_log_likelihood(ye)
            L, R = np.array(losses), np.array(rl)
            res = {
                "precision": precision, "arithmetic": {0: "fp32", 1: "tf32", 2: "bf16"}[arith], "world": world, "ema": ema,
                "steps": steps, "global_batch": GB, "variable_groups": groups,
                "p2p": os.environ.get("PGMVAE_P2P", "default"), "buckets": os.environ.get("PGMVAE_DP_BUCKETS", "0"),
                "loss_rel": float(np.abs(L[:, :3] - R[:, :3]).max() / np.abs(R[:, :3]).max()),
                "vq_loss_rel": float((np.abs(L[:, 3] - R[:, 3]) / np.maximum(np.abs(R[:, 3]), 1e-30)).max()),
                "sharded_exchange": sharded, "replicas_identical": len(set(digests)) == 1,
                "weight_rel": max(_rel(got[n], ref._get_tensor(n)) for n in tensors),
                "moment_rel": max(_rel(got[n], ref._get_tensor(n)) for n in moments),
                "codebook_rel": _rel(emb, ref._get_tensor("vq.embeddings")),
                "count_l1": float(np.abs(n1 - r1).sum() / max(r1.sum(), 1.0)),
                "count_total_ok": bool((n1 + n0).sum() == len(ye) * V),
                "pll_rel_sample_sharded": abs(pll_s - rpll) / abs(rpll),
                "pll_rel_variable_sharded": abs(pll_v - rpll) / abs(rpll),
            }
            del ref
        if world > 1:
            tdist.barrier()
        return res
    finally:
        if old is not None:
            ctx.set_precision(old)


def check(res, tol=1e-3):
    bad = [k for k in ("loss_rel", "weight_rel", "moment_rel", "count_l1", "pll_rel_sample_sharded", "pll_rel_variable_sharded")
           if not res[k] <= tol]
    if not res["replicas_identical"]:
        bad.append("replicas_identical")
    # a latent that sits within reduction-order rounding of a decision boundary may take the other code on one side,
    # which moves one code vector by ~(1 - decay) / (its size): a looser bar for the codebook maximum
    if not res["codebook_rel"] <= 10 * tol:
        bad.append("codebook_rel")
    if not res["count_total_ok"]:
        bad.append("count_total_ok")
    return bad


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _ffi.get_context(local)
    comm, rank, world = dist.init_from_env(ctx)
    cfg2 = ([50, 40, 30, 20], 69, 16, 128, 1024 // max(world, 1))
    wide = ([400, 200, 100, 50], 24, 64, 512, 96)            # cfg3-like widths: never on the chain kernels
    cases = []
    for prec in ("fp32", "tf32", "bf16"):
        for p2p in (("0", "1") if world <= 8 else ("0",)):
            for buckets in (("0", "1") if prec == "tf32" else ("0",)):
                cases.append((cfg2, prec, True, p2p, buckets, None))
    cases.append((cfg2, "tf32", False, "0", "0", None))
    cases.append((cfg2, "bf16", False, "0", "0", None))
    for prec in ("tf32", "bf16"):
        cases.append((wide, prec, True, "0", "0", "7"))      # several variable groups: per-group overlapped NCCL exchange
    cases.append((wide, "bf16", False, "0", "0", "5"))
    if world <= 8:
        # the same through the sharded peer-to-peer exchange fused with Adam (the default where the ranks can map
        # each other): "d" = environment untouched
        for prec in ("tf32", "bf16"):
            cases.append((wide, prec, True, "d", "0", "7"))
        cases.append((wide, "bf16", False, "d", "0", "5"))
    failed = 0
    for (units, V, D, K, prb), prec, ema, p2p, buckets, gv in cases:
        if p2p == "d":
            os.environ.pop("PGMVAE_P2P", None)
        else:
            os.environ["PGMVAE_P2P"] = p2p
        if buckets == "1":
            os.environ["PGMVAE_DP_BUCKETS"] = "1"
        else:
            os.environ.pop("PGMVAE_DP_BUCKETS", None)
        if gv:
            os.environ["PGMVAE_GROUP_VARS"] = gv
        else:
            os.environ.pop("PGMVAE_GROUP_VARS", None)
        if prec == "bf16":                                   # the bf16 layer-by-layer path also where the chains would fit
            os.environ["PGMVAE_NO_CHAIN"] = "1"
        else:
            os.environ.pop("PGMVAE_NO_CHAIN", None)
        res = dp_parity(ctx, comm, rank, world, units, V, D, K, prb, prec, ema=ema, local=local)
        if rank == 0:
            bad = check(res)
            failed += bool(bad)
            print(("FAIL " + ",".join(bad) if bad else "ok  ") + " " + " ".join(
                f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in res.items()), flush=True)
    if rank == 0:
        print(f"dp_check: {len(cases) - failed}/{len(cases)} cases within 1e-3", flush=True)
    if comm is not None:
        import torch.distributed as tdist
        tdist.barrier()
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
