"""Data-parallel parity check, run under torchrun on N GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        pgm-vae_b200/tools/dp_check.py
Every rank trains on its share of each global batch (NCCL allreduce of gradients, EMA statistics and loss
accumulators inside libpgmvae.so); rank 0 additionally trains a single-GPU model on the full batches and the two
must agree (same weights, losses, codebook, PLL counts)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgmvae import _ffi, data, dist  # noqa: E402
from core.model import VqVAE, Adam  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _ffi.get_context(local)
    comm, rank, world = dist.init_from_env(ctx)
    units, V, D, K, GB, steps = [50, 40, 30, 20], 69, 16, 128, 1024, 3
    y = data.synthetic_binary(GB * steps, V, seed=11)
    for ema in (True, False):
        m = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, seed=5, max_batch=GB, device=local, comm=comm)
        m.compile(optimizer=Adam(lr=1e-3))
        losses = []
        for s in range(steps):
            gb = y[s * GB:(s + 1) * GB]
            lo, hi = dist.shard_bounds(GB, rank, world)
            losses.append(m.train_on_batch(np.ascontiguousarray(gb[lo:hi]), global_batch=GB)["loss"])
        lo, hi = dist.shard_bounds(len(y), rank, world)
        n1, n0 = m.count(y[lo:hi])
        if rank == 0:
            ref = VqVAE(units, V, D, K, cost=0.25, decay=0.99, ema=ema, seed=5, max_batch=GB, device=local, comm=None)
            ref.compile(optimizer=Adam(lr=1e-3))
            rl = [ref.train_on_batch(np.ascontiguousarray(y[s * GB:(s + 1) * GB]))["loss"] for s in range(steps)]
            np.testing.assert_allclose(losses, rl, rtol=1e-5)
            for n in ("fd0.kernel", "fd4.bias", "fd9.kernel", "vq.embeddings"):
                a, b = m._get_tensor(n), ref._get_tensor(n)
                assert np.abs(a - b).max() <= 1e-3 * np.abs(b).max() + 1e-5, n
            r1, r0 = ref.count(y)
            assert np.abs(n1 - r1).sum() <= 0.001 * r1.sum(), "PLL counts differ"
            assert (n1 + n0).sum() == len(y) * V
            print(f"dp_check ema={ema}: world={world} losses {losses} == single-GPU {rl}: OK", flush=True)
    if comm is not None:
        import torch.distributed as tdist
        tdist.barrier()


if __name__ == "__main__":
    main()
