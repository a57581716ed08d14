"""VqVAE on B200 -- host mirror of the reference ``core/model.py`` (VqVAE, :14-148).

Same constructor ``VqVAE(units, nvar, dim, k, cost=0.5, decay=0.99, ema=True)``, same
attributes (``fd0..fd9``, ``vq_layer``, ``dist``) and methods (``__call__``, ``compile``,
``fit``, ``count``, ``cpt``, ``pseudo_log_likelihood``, ``get_probability``).  All state
lives in HBM inside a ``pgmvae_model`` handle of libpgmvae.so; one Keras fit step
(run.py:62) is one ``pgmvae_model_train_step`` call and ``count`` (core/model.py:58-82) is
one ``pgmvae_model_count`` call.  No TensorFlow, no CPU fallback.

Inputs: the reference feeds the materialised leave-one-out tensor ``x [N, V, V-1]``
(run.py:48-50).  Every method here accepts that tensor *or* the raw data matrix
``y [N, V]`` (0/1) from which it is built; the kernels only ever read ``y``.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

from pgmvae import _ffi
from core.dense import FatDense
from core.quantizer import VectorQuantizer, VectorQuantizerEMA


class Adam:
    """Stand-in for ``tf.keras.optimizers.Adam(lr=...)`` (run.py:60): carries the learning
    rate; the update itself is the fused Adam kernel (Keras form, eps=1e-7 added to sqrt(v))."""

    def __init__(self, lr=0.001, learning_rate=None, **kwargs):
        self.learning_rate = float(learning_rate if learning_rate is not None else lr)
        self.lr = self.learning_rate


class History:
    def __init__(self):
        self.history: Dict[str, List[float]] = {"loss": [], "mae": []}
        self.epoch: List[int] = []


def to_y(x) -> np.ndarray:
    """uint8 data matrix [N,V] from either y [N,V] or the reference's x [N,V,V-1]
    (row v of a sample = y without element v, run.py:46-50)."""
    a = x.numpy() if isinstance(x, _ffi.DeviceArray) else x
    if hasattr(a, "detach") and hasattr(a, "cpu"):
        a = a.detach().cpu().numpy()
    a = np.asarray(a)
    if a.ndim == 3:
        n, v, vm1 = a.shape
        if vm1 != v - 1:
            raise ValueError(f"expected leave-one-out inputs [N,V,V-1], got {a.shape}")
        y = np.empty((n, v), dtype=a.dtype)
        y[:, 0] = a[:, 1, 0]           # row 1 drops element 1, so it starts with y0
        y[:, 1:] = a[:, 0, :]          # row 0 drops element 0
        a = y
    if a.ndim != 2:
        raise ValueError(f"expected y [N,V] or x [N,V,V-1], got shape {a.shape}")
    if a.dtype != np.uint8:
        a = (a != 0).astype(np.uint8)
    return np.ascontiguousarray(a)


class VqVAE:
    """Many independent per-variable auto-encoders packed in one network around a
    per-variable VQ codebook (reference core/model.py:14-55)."""

    def __init__(self, units, nvar, dim, k, cost=0.5, decay=0.99, ema=True, *, seed=0, max_batch=1024,
                 device: Optional[int] = None, comm=None):
        if len(units) != 4:
            raise ValueError("units must list the 4 hidden widths (core/model.py:21-24 indexes units[0..3])")
        self.name = "vq_vae"
        self.units, self.nvar, self.dim, self.k = [int(u) for u in units], int(nvar), int(dim), int(k)
        self.cost, self.decay, self.ema, self.seed = float(cost), float(decay), bool(ema), int(seed)
        self.epsilon = 1e-5
        self.ctx = _ffi.get_context(device)
        self.comm = comm
        self._h = C.c_void_p()
        self.max_batch = 0
        self._create(int(max_batch))
        _ffi.check(_ffi.lib().pgmvae_model_init(self._h, C.c_uint64(self.seed)))
        act, init = "selu", "he_uniform"
        widths = self.units + [self.dim] + self.units[::-1]
        self._layers = []
        for i in range(9):
            self._layers.append(FatDense(widths[i], activation=act, kernel_initializer=init))
        self._layers.append(FatDense(self.nvar - 1, activation="sigmoid", kernel_initializer="glorot_uniform"))
        for i, l in enumerate(self._layers):
            l._bind(self, i)
            setattr(self, f"fd{i}", l)
        if self.ema:
            self.vq_layer = VectorQuantizerEMA(embedding_dim=dim, num_embeddings=k, commitment_cost=cost, decay=decay,
                                               num_var=nvar)
        else:
            self.vq_layer = VectorQuantizer(embedding_dim=dim, num_embeddings=k, commitment_cost=cost, num_var=nvar)
        self.vq_layer._bind(self)
        self.dist = np.zeros((self.nvar, self.k), dtype=np.float64)        # core/model.py:37
        self.losses: List[float] = []
        self.optimizer: Optional[Adam] = None
        self._adam_t = 0
        self._ema_steps = 0

    # ---- handle management -------------------------------------------------------
    def _create(self, max_batch: int):
        units = (C.c_int * 4)(*self.units)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().pgmvae_model_create(self.ctx.h, units, self.nvar, self.dim, self.k, self.cost, self.decay,
                                                  self.epsilon, int(self.ema), max_batch, C.byref(h)))
        self._h, self.max_batch = h, max_batch
        self._setup_p2p()

    def _setup_p2p(self):
        """Data parallel on one node: map every rank's gradient buffer into every other rank (CUDA IPC) so that the
        library can fuse the gradient exchange with Adam (pgmvae_model_p2p_*).  Collective: every rank creates its
        model at the same point.  torch.distributed (gloo) only carries the 384 handle bytes."""
        comm = self.comm
        if comm is None or getattr(comm, "h", None) is None or comm.nranks < 2 or comm.nranks > 8:
            return
        # Which exchanges use the mapping is the library's decision (pgmvae_model_p2p_import): narrow models (chain kernels)
        # sum the whole gradient buffer peer-to-peer at two ranks only (measured, cfg2 step: 0.686 vs 0.702 ms NCCL at 2
        # GPUs, 0.80 vs 0.77 ms at 8, where every rank reads all eight buffers); wide models (per-group path) run the
        # SHARDED exchange fused with Adam at any rank count.  PGMVAE_P2P=0: no mapping, NCCL everywhere.
        if os.environ.get("PGMVAE_P2P") == "0":
            return
        import socket
        import torch.distributed as dist
        lib = _ffi.lib()
        nbytes = 384
        # every rank must sit on the same host with peer access between all devices; otherwise (two nodes, GPUs
        # without NVLink / PCIe peer access) the NCCL path is the one that works.  The decision is collective.
        infos = [None] * comm.nranks
        dist.all_gather_object(infos, (socket.gethostname(), int(self.ctx.device)))
        ok = len({h for h, _ in infos}) == 1
        if ok:
            can = C.c_int(0)
            for _, dev in infos:
                if dev != int(self.ctx.device):
                    _ffi.check(lib.pgmvae_device_can_access_peer(int(self.ctx.device), int(dev), C.byref(can)))
                    ok = ok and bool(can.value)
        flags = [None] * comm.nranks
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            return
        buf = C.create_string_buffer(nbytes)
        rc = lib.pgmvae_model_p2p_export(self._h, buf)
        handles = [None] * comm.nranks
        dist.all_gather_object(handles, buf.raw if rc == 0 else None)
        good = all(h is not None for h in handles)
        if good:
            blob = C.create_string_buffer(b"".join(handles), nbytes * comm.nranks)
            rc = lib.pgmvae_model_p2p_import(self._h, comm.rank, comm.nranks, blob)
            good = rc == 0
        flags = [None] * comm.nranks
        dist.all_gather_object(flags, bool(good))
        if not all(flags):
            # some rank could not map its peers: nobody uses the peer-to-peer exchange
            _ffi.check(lib.pgmvae_model_p2p_disable(self._h))
        dist.barrier()

    def _ensure_capacity(self, batch: int, collective: bool = False):
        """The workspace is sized for max_batch samples; grow it (state is carried over).  Re-creating the model is
        a COLLECTIVE operation under data parallelism (the peer-to-peer mapping is set up again), so there it only
        happens where every rank arrives with the same number (``collective=True``: fit / train_on_batch size the
        workspace for ceil(global_batch / world), the largest share any rank can get); rank-local calls
        (``model(x)``, evaluation) must stay within the capacity."""
        if batch <= self.max_batch:
            return
        multi = self.comm is not None and getattr(self.comm, "nranks", 1) > 1 and getattr(self.comm, "h", None) is not None
        if multi and not collective:
            raise ValueError(f"batch of {batch} samples exceeds the model's capacity ({self.max_batch}); on a data-parallel "
                             f"model the workspace only grows inside fit()/train_on_batch(), where every rank takes part")
        state = self.state_dict()
        old = self._h
        self._create(int(batch))
        _ffi.lib().pgmvae_model_destroy(old)
        self.load_state_dict(state)

    def tensor_names(self) -> List[str]:
        names = [f"fd{i}.{s}" for i in range(10) for s in ("kernel", "bias")] + ["vq.embeddings"]
        if self.ema:
            names += ["vq.ema_w", "vq.ema_cluster_size", "vq.biased_w", "vq.biased_c"]
        return names

    def _tensor_shape(self, name: str):
        V, D, K = self.nvar, self.dim, self.k
        base = name.split(".", 1)[1] if name.split(".", 1)[0] in ("grad", "adam_m", "adam_v") else name
        if base.startswith("fd"):
            i = int(base[2])
            chain = [V - 1] + self.units + [D] + self.units[::-1] + [V - 1]
            return (V, chain[i], chain[i + 1]) if base.endswith("kernel") else (V, 1, chain[i + 1])
        if base in ("vq.embeddings", "vq.ema_w", "vq.biased_w", "vq.stat_w"):
            return (V, D, K)
        return (V, K)

    def _get_tensor(self, name: str) -> np.ndarray:
        out = np.empty(self._tensor_shape(name), dtype=np.float32)
        _ffi.check(_ffi.lib().pgmvae_model_get_tensor(self._h, name.encode(), out.ctypes.data, out.size))
        return out

    def _set_tensor(self, name: str, value):
        v = np.ascontiguousarray(value, dtype=np.float32)
        if v.shape != self._tensor_shape(name):
            raise ValueError(f"{name}: expected shape {self._tensor_shape(name)}, got {v.shape}")
        _ffi.check(_ffi.lib().pgmvae_model_set_tensor(self._h, name.encode(), v.ctypes.data, v.size))

    def state_dict(self) -> Dict[str, np.ndarray]:
        """Every tensor in the reference layouts + the Adam moments.  After data-parallel train_on_batch() steps with the
        sharded peer-to-peer exchange the moments and the fp32 master kernels live with the rank that owns a shard: the
        gather that completes them (sync_state) is COLLECTIVE, so there state_dict() / save_weights() must be called on
        every rank.  fit() ends with that gather."""
        self.sync_state()
        sd = {n: self._get_tensor(n) for n in self.tensor_names()}
        train = [f"fd{i}.{s}" for i in range(10) for s in ("kernel", "bias")] + ([] if self.ema else ["vq.embeddings"])
        for n in train:
            sd["adam_m." + n] = self._get_tensor("adam_m." + n)
            sd["adam_v." + n] = self._get_tensor("adam_v." + n)
        sd["_adam_t"] = np.int64(self._adam_t)
        sd["_ema_steps"] = np.int64(self._ema_steps)
        return sd

    def sync_state(self):
        """COLLECTIVE under data parallelism with the sharded exchange (no-op otherwise): complete the optimiser state
        and the fp32 master weights on every rank (pgmvae_model_p2p_sync_state)."""
        lib = _ffi.lib()
        if lib.pgmvae_model_p2p_state_sharded(self._h):
            _ffi.check(lib.pgmvae_model_p2p_sync_state(self._h))

    def load_state_dict(self, sd: Dict[str, np.ndarray]):
        for n, v in sd.items():
            if not n.startswith("_"):
                self._set_tensor(n, v)
        if "_adam_t" in sd:
            self._adam_t = int(sd["_adam_t"])
            _ffi.check(_ffi.lib().pgmvae_model_set_adam_step(self._h, self._adam_t))
        if "_ema_steps" in sd:
            self._ema_steps = int(sd["_ema_steps"])
            _ffi.check(_ffi.lib().pgmvae_model_set_ema_steps(self._h, self._ema_steps, self._ema_steps))

    def set_weights_from(self, params: Dict[str, np.ndarray]):
        """Inject reference-layout weights ({'fd0.kernel':..., 'vq.embeddings':...}); the EMA
        shadow ``ema_w`` is re-initialised from the codebook (core/quantizer.py:117)."""
        for n, v in params.items():
            self._set_tensor(n, np.asarray(v, dtype=np.float32))
        if self.ema and "vq.embeddings" in params and "vq.ema_w" not in params:
            self._set_tensor("vq.ema_w", np.asarray(params["vq.embeddings"], dtype=np.float32))

    # ---- forward -----------------------------------------------------------------
    def __call__(self, inputs, training=None, code_only=False, fts=None):
        """reference core/model.py:39-55."""
        if fts is not None:
            return self._call_fts(inputs, code_only, fts)
        y = to_y(inputs)
        B = y.shape[0]
        self._ensure_capacity(B)
        L = _ffi.lib()
        if code_only:
            self.losses = [0.0]
            idx = self.encode(y)
            if idx.size * self.k * 4 > (2 << 30):
                raise MemoryError(f"code_only=True returns the reference's one-hot tensor [V,B,K] = {idx.shape + (self.k,)} float32 "
                                  f"({idx.size * self.k * 4 / 2**30:.1f} GiB); use VqVAE.encode(y) for the codes [V,B] (count / cpt / "
                                  f"pseudo_log_likelihood never build the one-hot tensor)")
            out = np.zeros(idx.shape + (self.k,), dtype=np.float32)           # one-hot [V,B,K] (core/quantizer.py:139,159)
            np.put_along_axis(out, idx[..., None].astype(np.int64), 1.0, axis=-1)
            return out
        Vp = (self.nvar + 7) // 8 * 8
        out = _ffi.DeviceArray(self.ctx, (self.nvar, self.max_batch, Vp), np.float32, zero=False)
        met = (C.c_double * 4)()
        _ffi.check(L.pgmvae_model_forward(self._h, y.ctypes.data, 0, B, 1 if training else 0, out.ptr, met))
        if training and self.ema:
            self._ema_steps += 1
        self.losses = [met[3]]
        o = out.numpy()[:, :B, :self.nvar]                                   # [V,B,V] expanded
        V = self.nvar
        keep = ~np.eye(V, dtype=bool)
        rec = o.transpose(1, 0, 2)[:, keep].reshape(B, V, V - 1)            # drop column v of net v
        return np.ascontiguousarray(rec)

    def encode(self, inputs) -> np.ndarray:
        """Codes of every variable's latent, int32 [V,B]: the encoder fd0..fd4 + VQ assignment (what code_only=True
        computes, core/model.py:48, without materialising the one-hot tensor)."""
        y = to_y(inputs)
        B = y.shape[0]
        self._ensure_capacity(B)
        idx = _ffi.DeviceArray(self.ctx, (self.nvar, B), np.int32, zero=False)
        _ffi.check(_ffi.lib().pgmvae_model_encode(self._h, y.ctypes.data, 0, B, idx.ptr))
        return idx.numpy()

    def _call_fts(self, inputs, code_only, fts):
        """Sub-net path (core/model.py:41, ``fts`` branches of every layer): inputs [F,B,V-1].  code_only (what
        ``get_probability`` and the Gibbs sampler use) runs on the device against the model's weights in place
        (``pgmvae_model_fts_encode``): one launch per layer for ALL selected networks, no weight copies."""
        fts = np.asarray(fts, dtype=np.int32).reshape(-1)
        if code_only:
            x = _ffi.as_host_f32(inputs.numpy() if isinstance(inputs, _ffi.DeviceArray) else inputs)
            F, B, vm1 = x.shape
            if vm1 != self.nvar - 1 or F != len(fts):
                raise ValueError(f"expected inputs [len(fts), B, V-1], got {x.shape}")
            # expanded over all V data columns: net v ignores column v (its weight row is zero), so any value will do there
            xe = np.zeros((F, B, self.nvar), np.float32)
            for f, v in enumerate(fts):
                xe[f, :, :v] = x[f, :, :v]
                xe[f, :, v + 1:] = x[f, :, v:]
            idx = np.empty((F, B), np.int32)
            _ffi.check(_ffi.lib().pgmvae_model_fts_encode(self._h, xe.ctypes.data, np.ascontiguousarray(fts).ctypes.data, F, B,
                                                          idx.ctypes.data))
            self.losses = [0.0]
            return idx
        x = inputs
        for i in range(5):
            x = self._layers[i](x, fts=fts)
        x = self.vq_layer(x, training=None, code_only=False, fts=fts)
        for i in range(5, 10):
            x = self._layers[i](x, fts=fts)
        return x.numpy()

    # ---- training ------------------------------------------------------------------
    def compile(self, optimizer=None, loss="mse", metrics=None):
        """run.py:61.  Only the reference configuration is implemented."""
        if loss not in ("mse", "mean_squared_error"):
            raise NotImplementedError("only loss='mse' (the reference configuration) is implemented")
        for m in (metrics or []):
            if m not in ("mae", "mean_absolute_error"):
                raise NotImplementedError("only metrics=['mae'] is implemented")
        if optimizer is None:
            optimizer = Adam()
        lr = getattr(optimizer, "learning_rate", getattr(optimizer, "lr", None))
        self.optimizer = optimizer if isinstance(optimizer, Adam) else Adam(lr=float(lr))

    def train_on_batch(self, y_batch: np.ndarray, global_batch: Optional[int] = None, sync: bool = True):
        """One Keras fit step on a uint8 batch [B,V]; returns {loss, mse, mae, vq_loss} if sync."""
        if self.optimizer is None:
            self.compile()
        B = y_batch.shape[0]
        world = self.comm.nranks if self.comm is not None else 1
        # the same number on every rank: the largest share of the global batch (shares differ by at most one sample)
        self._ensure_capacity(max(B, -(-int(global_batch or B) // world)) if world > 1 else B, collective=True)
        met = (C.c_double * 4)() if sync else None
        comm_h = self.comm.h if self.comm is not None else None
        _ffi.check(_ffi.lib().pgmvae_model_train_step(self._h, y_batch.ctypes.data, 0, B, int(global_batch or B),
                                                      self.optimizer.learning_rate, comm_h, 0, met))
        self._adam_t += 1
        if self.ema:
            self._ema_steps += 1
        if sync:
            return {"loss": met[0], "mse": met[1], "mae": met[2], "vq_loss": met[3]}
        return None

    def fit(self, x, y=None, batch_size=32, epochs=1, callbacks=None, verbose=0, shuffle=True, order=None):
        """``model.fit(train_x, train_x, batch_size, epochs)`` (run.py:62): per-epoch reshuffle,
        last partial batch kept, loss/mae reported as sample-weighted epoch means.

        ``order`` (list of per-epoch permutations) pins the batch order for parity runs; by
        default ``np.random.permutation`` is used, which run.py seeds (run.py:36).  With a
        data-parallel ``comm`` every rank must pass the same order; each rank then takes its
        contiguous share of every global batch."""
        if self.optimizer is None:
            self.compile()
        data = to_y(x)
        n = data.shape[0]
        rank, world = (self.comm.rank, self.comm.nranks) if self.comm is not None else (0, 1)
        hist = History()
        for ep in range(int(epochs)):
            if order is not None:
                perm = np.asarray(order[ep])
            elif shuffle:
                perm = np.random.permutation(n)
            else:
                perm = np.arange(n)
            t0 = time.time()
            sums = np.zeros(2)
            seen = 0
            for s in range(0, n, batch_size):
                gidx = perm[s:s + batch_size]
                gb = len(gidx)
                lo, hi = gb * rank // world, gb * (rank + 1) // world
                if hi <= lo:
                    raise ValueError("a data-parallel rank received an empty share of a batch")
                batch = np.ascontiguousarray(data[gidx[lo:hi]])
                m = self.train_on_batch(batch, global_batch=gb, sync=True)
                sums += np.array([m["loss"], m["mae"]]) * gb
                seen += gb
            hist.epoch.append(ep)
            hist.history["loss"].append(sums[0] / seen)
            hist.history["mae"].append(sums[1] / seen)
            if verbose and rank == 0:
                print(f"Epoch {ep + 1}/{epochs} - {time.time() - t0:.2f}s - loss: {sums[0] / seen:.6f} "
                      f"- mae: {sums[1] / seen:.6f}", flush=True)
            for cb in (callbacks or []):
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(ep, {"loss": sums[0] / seen, "mae": sums[1] / seen})
        self.sync_state()      # (data parallel, sharded exchange: every rank leaves fit() with the complete model)
        return hist

    # ---- stage 2 ---------------------------------------------------------------------
    def var_shard(self):
        """[v0, v1): the variables this rank owns when stage 2 is sharded over variable groups."""
        if self.comm is None or self.comm.nranks <= 1:
            return 0, self.nvar
        return self.nvar * self.comm.rank // self.comm.nranks, self.nvar * (self.comm.rank + 1) // self.comm.nranks

    def count(self, x, y=None, var_range=None):
        """n1[v,k] = #(y_v = 1, code_v = k), n0 likewise (reference core/model.py:58-82),
        as float64 [V,K].  With a communicator the samples are this rank's shard and the
        counts are summed over ranks.  ``var_range=(v0, v1)``: only those variables are evaluated (on all the
        samples given) and nothing is exchanged -- rows outside the range stay zero."""
        data = to_y(y if y is not None else x)
        n = data.shape[0]          # pgmvae_model_count walks the samples in chunks of max_batch
        n1 = np.zeros((self.nvar, self.k), dtype=np.uint64)
        n0 = np.zeros((self.nvar, self.k), dtype=np.uint64)
        if var_range is not None:
            v0, v1 = int(var_range[0]), int(var_range[1])
            _ffi.check(_ffi.lib().pgmvae_model_count_vars(self._h, data.ctypes.data, 0, n, v0, v1, n1.ctypes.data,
                                                          n0.ctypes.data))
            return n1.astype(np.float64), n0.astype(np.float64)
        _ffi.check(_ffi.lib().pgmvae_model_count(self._h, data.ctypes.data, 0, n, n1.ctypes.data, n0.ctypes.data))
        if self.comm is not None and self.comm.nranks > 1:
            n1, n0 = self.comm.allreduce_u64(n1), self.comm.allreduce_u64(n0)
        return n1.astype(np.float64), n0.astype(np.float64)

    def count_stream(self, chunks, rows_per_chunk: int = 32768, var_range=None):
        """``count`` over an iterator of chunks y [n_i, V] (``pgmvae.data.iter_binary_csv`` / ``iter_array`` / any
        generator) for data sets that fit neither host nor device memory (run.py:53, the author's TODO).  A background
        thread reads / parses the next chunks into pinned host buffers while the device works on the current one;
        counts accumulate on the device and come back once.  Returns (n1, n0, number of samples)."""
        from pgmvae.data import PinnedPrefetcher
        v0, v1 = (0, self.nvar) if var_range is None else (int(var_range[0]), int(var_range[1]))
        L = _ffi.lib()
        n1 = np.zeros((self.nvar, self.k), dtype=np.uint64)
        n0 = np.zeros((self.nvar, self.k), dtype=np.uint64)
        pf = PinnedPrefetcher(self.ctx, chunks, int(rows_per_chunk), self.nvar)
        total, in_flight = 0, []
        try:
            _ffi.check(L.pgmvae_model_count_begin(self._h))
            for buf, rows in pf:
                _ffi.check(L.pgmvae_model_count_add(self._h, buf.ctypes.data, 0, rows, v0, v1))      # asynchronous
                total += rows
                in_flight.append(buf)
                if len(in_flight) >= 2:              # the chunk before the one just enqueued has been consumed
                    self.ctx.sync()
                    for b in in_flight:
                        pf.release(b)
                    in_flight = []
            _ffi.check(L.pgmvae_model_count_end(self._h, v0, v1, n1.ctypes.data, n0.ctypes.data))
        finally:
            self.ctx.sync()
            pf.close()
        if var_range is None and self.comm is not None and self.comm.nranks > 1:
            n1, n0 = self.comm.allreduce_u64(n1), self.comm.allreduce_u64(n0)
            total = int(self.comm.allreduce_u64(np.array([total], np.uint64))[0])
        return n1.astype(np.float64), n0.astype(np.float64), total

    def pseudo_log_likelihood_stream(self, chunks, rows_per_chunk: int = 32768):
        """``pseudo_log_likelihood`` over an iterator of chunks (see ``count_stream``); ``self.dist`` must hold the CPT."""
        n1, n0, n = self.count_stream(chunks, rows_per_chunk)
        ctx, L = self.ctx, _ffi.lib()
        d1 = _ffi.DeviceArray.from_numpy(ctx, n1.astype(np.uint64))
        d0 = _ffi.DeviceArray.from_numpy(ctx, n0.astype(np.uint64))
        dd = _ffi.DeviceArray.from_numpy(ctx, np.ascontiguousarray(self.dist, dtype=np.float64))
        out = _ffi.DeviceArray(ctx, (1,), np.float64)
        _ffi.check(L.pgmvae_pll_reduce(ctx.h, None, d1.ptr, d0.ptr, dd.ptr, self.nvar * self.k, out.ptr))
        return float(out.numpy()[0]) / max(n, 1)

    def cpt(self, x, y=None, shard="samples"):
        """p(y=1 | code=k) with additive smoothing (reference core/model.py:85-88).  shard="variables": every rank
        passes ALL samples and fills only the rows of its own variables (``var_shard``)."""
        n1, n0 = self.count(x, y, var_range=self.var_shard() if shard == "variables" else None)
        return (n1 + 0.8) / (n1 + n0 + 1.6)

    def pseudo_log_likelihood(self, x, y=None, total: Optional[int] = None, shard="samples"):
        """Average pseudo log-likelihood (reference core/model.py:91-96); the float64
        reduction runs on the device (pgmvae_pll_reduce).

        shard="samples" (default): x is this rank's shard of the samples, the [V,K] counts are summed over ranks.
        shard="variables": x holds ALL samples on every rank; each rank evaluates only its own variables
        (``var_shard``) against its rows of ``self.dist`` and ONE float64 scalar is all-reduced."""
        data = to_y(y if y is not None else x)
        by_var = shard == "variables"
        v0, v1 = self.var_shard() if by_var else (0, self.nvar)
        n1, n0 = self.count(data, var_range=(v0, v1) if by_var else None)
        n = int(total if total is not None else data.shape[0])
        ctx, L = self.ctx, _ffi.lib()
        sl = slice(v0, v1)
        d1 = _ffi.DeviceArray.from_numpy(ctx, np.ascontiguousarray(n1[sl]).astype(np.uint64))
        d0 = _ffi.DeviceArray.from_numpy(ctx, np.ascontiguousarray(n0[sl]).astype(np.uint64))
        dd = _ffi.DeviceArray.from_numpy(ctx, np.ascontiguousarray(self.dist[sl], dtype=np.float64))
        out = _ffi.DeviceArray(ctx, (1,), np.float64)
        if v1 > v0:
            _ffi.check(L.pgmvae_pll_reduce(ctx.h, None, d1.ptr, d0.ptr, dd.ptr, (v1 - v0) * self.k, out.ptr))
        s = float(out.numpy()[0])
        if by_var and self.comm is not None and self.comm.nranks > 1:
            s = float(self.comm.allreduce_f64(np.array([s], np.float64))[0])
        return s / n

    def get_probability(self, x, fts=None):
        """p(y_i = 1 | code) of the selected nets (reference core/model.py:99-108);
        x [F, B, V-1], returns [F, B] float32."""
        fts = np.asarray(fts, dtype=np.int64).reshape(-1)
        enc_idx = self(x, code_only=True, fts=fts)                          # [F,B]
        prb = self.dist[fts].astype(np.float32)
        return np.take_along_axis(prb, enc_idx, axis=1)

    def conditional_marginal_log_likelihood(self, x, p1, num_smp, burn_in, verbose=False, uniform=None, seed=None):
        """Conditional marginal log-likelihood by block Gibbs sampling (reference core/model.py:110-148; intended
        call at run.py:74: ``p1=n_var//10, num_smp=3000, burn_in=150``).  The sampler runs ON THE DEVICE
        (``pgmvae_model_gibbs_cmll``): state, counters and the sub-net evaluation of every step stay in HBM, the host only
        enqueues launches.  ``uniform(shape)`` injects the U[0,1) draws step by step (parity runs: the reference's
        tf.random stream cannot be reproduced without TensorFlow); otherwise a counter-based generator on the device
        is seeded with ``seed`` (default: drawn from numpy's global generator, which run.py seeds)."""
        data = to_y(x)
        B, V = data.shape
        blocks = -(-V // int(p1))
        steps = int(num_smp) * int(p1)
        uni = None
        if uniform is not None:
            uni = np.ascontiguousarray(np.stack([np.asarray(uniform((blocks, B)), dtype=np.float32) for _ in range(steps)]))
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        out = C.c_double(0.0)
        dist = np.ascontiguousarray(self.dist, dtype=np.float64)
        _ffi.check(_ffi.lib().pgmvae_model_gibbs_cmll(self._h, data.ctypes.data, B, int(p1), int(num_smp), int(burn_in),
                                                      dist.ctypes.data, C.c_uint64(seed),
                                                      uni.ctypes.data if uni is not None else None, C.byref(out)))
        return float(out.value)

    def device_bytes(self) -> int:
        return int(_ffi.lib().pgmvae_model_device_bytes(self._h))

    def group_size(self) -> int:
        """Variables per workspace group (the step walks the V independent networks group by group)."""
        return int(_ffi.lib().pgmvae_model_group_size(self._h))

    @staticmethod
    def _npz_path(path: str) -> str:
        path = os.fspath(path)
        return path if path.endswith(".npz") else path + ".npz"          # np.savez appends the suffix itself

    def save_weights(self, path: str):
        """Every tensor in the reference's layouts ([V,in,units], [V,1,units], [V,D,K]) + optimiser / EMA state
        (run.py:63 intent).  tools/tf_crosscheck.py loads such a file into the unmodified reference."""
        np.savez(self._npz_path(path), **self.state_dict())

    def load_weights(self, path: str):
        with np.load(self._npz_path(path)) as z:
            self.load_state_dict({k: z[k] for k in z.files})

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().pgmvae_model_destroy(self._h)
        except Exception:
            pass


def to_y_float(x):
    """[B,V] data as float32 (accepts uint8 / float arrays)."""
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)
