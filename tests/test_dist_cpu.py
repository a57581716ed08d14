"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, global
normalisation of the sharded gradients, EMA-statistic and PLL-count reductions, and the sharded
exchange (owner-computes reduce-scatter + Adam + all-gather, ownership from the library).  Compute is
the oracle (this file is test code); the reductions go through pgmvae.dist.GlooComm, which
has the interface of the NCCL communicator the product path uses."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "pgm-vae_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import pgmvae_oracle as O
    import np_expanded as NE
    from pgmvae.dist import GlooComm, shard_bounds
    comm = GlooComm()
    units, V, D, K, B = [6, 5, 4, 3], 5, 2, 4, 22
    params = {k: v.numpy() for k, v in O.init_params(units, V, D, K, seed=3).items()}
    y = O.synthetic_binary(B, V, seed=1)
    lo, hi = shard_bounds(B, rank, world)
    # each rank: gradients of its share with GLOBAL normalisation, then sum over ranks
    met, g_local, aux = NE.step_grads(params, y[lo:hi], D, K, 0.25, True, global_B=B)
    g_sum = {n: comm.allreduce_f64(g) for n, g in g_local.items()}
    m_sum = comm.allreduce_f64(np.array([met["mse"], met["mae"], met["vq_loss"]]))
    # EMA statistics of the shard, reduced before the update
    c, dw = O.ema_stats(aux["z"].astype(np.float32), aux["idx"], K)
    c_sum, dw_sum = comm.allreduce_f32(c.numpy()), comm.allreduce_f32(dw.numpy())
    # PLL counts of the shard
    n1 = np.zeros((V, K), np.uint64)
    for v in range(V):
        np.add.at(n1[v], aux["idx"][v][y[lo:hi, v] != 0], 1)
    n1_sum = comm.allreduce_u64(n1)
    if rank == 0:
        met_f, g_full, aux_f = NE.step_grads(params, y, D, K, 0.25, True)
        for n in g_full:
            np.testing.assert_allclose(g_sum[n], g_full[n], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(m_sum, [met_f["mse"], met_f["mae"], met_f["vq_loss"]], rtol=1e-9)
        cf, dwf = O.ema_stats(aux_f["z"].astype(np.float32), aux_f["idx"], K)
        np.testing.assert_array_equal(c_sum, cf.numpy())
        np.testing.assert_allclose(dw_sum, dwf.numpy(), rtol=1e-5, atol=1e-6)
        n1f = np.zeros((V, K), np.uint64)
        for v in range(V):
            np.add.at(n1f[v], aux_f["idx"][v][y[:, v] != 0], 1)
        np.testing.assert_array_equal(n1_sum, n1f)
        out.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def _adam_f32(p, m, v, g, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """Keras-form Adam in fp32, the expression of csrc/pll.cu: adam_kernel / csrc/model.cu: p2p_shard_adam_kernel"""
    f = np.float32
    alpha = f(lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    m = m + (g - m) * f(1.0 - b1)
    v = v + (g * g - v) * f(1.0 - b2)
    return p - (m * alpha) / (np.sqrt(v) + f(eps)), m, v


def _worker_sharded(rank, world, port, out):
    """The SHARDED exchange of csrc/model.cu (p2p_shard_adam_kernel) restated on CPU over gloo: per variable group the
    owner of a variable (the library's own pgmvae_p2p_shard_bounds) sums the ranks' partial gradients in rank order,
    applies Adam to its shard and hands the new values to everybody; moments stay with the owner until the gather.  Must
    equal all-reduce + replicated Adam, leave the replicas bit-identical, and complete the moments after the gather."""
    import ctypes as C
    import hashlib
    for p in (os.path.join(ROOT, "pgm-vae_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import pgmvae_oracle as O
    import np_expanded as NE
    from pgmvae import _ffi
    from pgmvae.dist import shard_bounds
    L = _ffi.load_library()
    units, V, D, K, B, Vg = [6, 5, 4, 3], 5, 2, 4, 22, 2           # variable groups {0,1} {2,3} {4}
    params = {k: v.numpy().astype(np.float32) for k, v in O.init_params(units, V, D, K, seed=3).items()}
    names = [n for n in params if n.startswith("fd")]              # the dense tensors ([V, ...]); EMA codebook: no Adam
    mom_m = {n: np.zeros_like(params[n]) for n in names}
    mom_v = {n: np.zeros_like(params[n]) for n in names}
    ref_p = {n: params[n].copy() for n in names}
    ref_m = {n: np.zeros_like(params[n]) for n in names}
    ref_v = {n: np.zeros_like(params[n]) for n in names}
    owner = np.full(V, -1)
    for g0 in range(0, V, Vg):
        for r in range(world):
            lo, hi = C.c_int(), C.c_int()
            assert L.pgmvae_p2p_shard_bounds(g0, min(Vg, V - g0), r, world, C.byref(lo), C.byref(hi)) == 0
            owner[lo.value:hi.value] = r
    assert (owner >= 0).all()
    mine = owner == rank
    for t in (1, 2):
        y = O.synthetic_binary(B, V, seed=10 + t)
        lo, hi = shard_bounds(B, rank, world)
        full = dict(params)
        _, g_local, _ = NE.step_grads(full, y[lo:hi], D, K, 0.25, True, global_B=B)
        g32 = {n: g_local[n].astype(np.float32) for n in names}
        gathered = [None] * world
        dist.all_gather_object(gathered, g32)                       # "every rank maps every rank's gradient buffer"
        for n in names:
            g = gathered[0][n].copy()
            for q in range(1, world):
                g = g + gathered[q][n]                              # rank order, fp32
            # replicated reference: every rank updates everything
            ref_p[n], ref_m[n], ref_v[n] = _adam_f32(ref_p[n], ref_m[n], ref_v[n], g, t)
            # sharded: only the owned variables
            pn, mn, vn = _adam_f32(params[n][mine], mom_m[n][mine], mom_v[n][mine], g[mine], t)
            params[n] = params[n].copy()
            params[n][mine], mom_m[n][mine], mom_v[n][mine] = pn, mn, vn
        handed = [None] * world
        dist.all_gather_object(handed, {n: params[n][mine] for n in names})      # "writes into the buffers of all ranks"
        for q in range(world):
            for n in names:
                params[n][owner == q] = handed[q][n]
    for n in names:
        np.testing.assert_array_equal(params[n], ref_p[n])          # same arithmetic, same order: bit-equal
        assert not np.array_equal(mom_m[n], ref_m[n]) or mine.all()    # the moments are sharded ...
    state = [None] * world
    dist.all_gather_object(state, {n: (mom_m[n][mine], mom_v[n][mine]) for n in names})   # ... until the gather (sync_state)
    for q in range(world):
        for n in names:
            mom_m[n][owner == q], mom_v[n][owner == q] = state[q][n]
    for n in names:
        np.testing.assert_array_equal(mom_m[n], ref_m[n])
        np.testing.assert_array_equal(mom_v[n], ref_v[n])
    digest = hashlib.sha1(b"".join(params[n].tobytes() for n in sorted(names))).hexdigest()
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    assert len(set(digests)) == 1
    if rank == 0:
        out.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_equals_replicated_adam_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_data_parallel_reductions_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_shard_bounds_partition():
    from pgmvae.dist import shard_bounds
    for n in (0, 1, 7, 4096, 16181):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_sharded_exchange_ownership_partition():
    """The sharded peer-to-peer exchange (csrc/model.cu: p2p_shard_adam_kernel) gives every variable of every variable
    group exactly one owner rank: the library's own ownership function, over the group walks the training step does
    (cfg3: 1556 variables in groups of 148, last group 76; fewer variables than ranks; one group)."""
    import ctypes as C
    for p in (os.path.join(ROOT, "pgm-vae_b200"),):
        if p not in sys.path:
            sys.path.insert(0, p)
    from pgmvae import _ffi
    L = _ffi.load_library()
    for V, Vg in ((1556, 148), (69, 20), (24, 7), (5, 5), (3, 148)):
        for R in (2, 3, 4, 8):
            owner = np.full(V, -1)
            for g0 in range(0, V, Vg):
                Gn = min(Vg, V - g0)
                sizes = []
                for r in range(R):
                    lo, hi = C.c_int(-1), C.c_int(-1)
                    assert L.pgmvae_p2p_shard_bounds(g0, Gn, r, R, C.byref(lo), C.byref(hi)) == 0
                    assert g0 <= lo.value <= hi.value <= g0 + Gn
                    assert (owner[lo.value:hi.value] == -1).all()
                    owner[lo.value:hi.value] = r
                    sizes.append(hi.value - lo.value)
                assert max(sizes) - min(sizes) <= 1
            assert (owner >= 0).all()
    lo, hi = C.c_int(0), C.c_int(0)
    assert L.pgmvae_p2p_shard_bounds(0, 4, 2, 2, C.byref(lo), C.byref(hi)) != 0       # rank out of range
