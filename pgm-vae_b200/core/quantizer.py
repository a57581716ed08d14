"""Vector-quantiser layers on B200 -- host mirror of the reference ``core/quantizer.py``.

``VectorQuantizer`` (core/quantizer.py:13-71) and ``VectorQuantizerEMA`` (:74-176) keep the
reference constructors, the ``embeddings [V, D, K]`` / ``ema_cluster_size [V, K]`` /
``ema_w [V, D, K]`` attributes, the ``(inputs, training, code_only, fts)`` call and the
``losses`` side channel (Keras ``add_loss``, :58,:161).  The TensorFlow ops are replaced by
library kernels: fused distance + argmin (``pgmvae_vq_assign``, never materialising
``[V,B,K]``), gather/loss/straight-through (``pgmvae_vq_quantize``), the segmented
scatter-add for the EMA statistics (``pgmvae_ema_stats``) and the debiased-EMA / Laplace /
normalise update (``pgmvae_ema_apply``, TF ``assign_moving_average(zero_debias=True)``).

In HBM the codebook is held code-major ``[V, K, D]``; the reference layout is produced on
attribute access.  A layer is standalone or bound to a ``core.model.VqVAE`` handle.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from pgmvae import _ffi
from core.dense import initialize


class _VQBase:
    ema = False

    def __init__(self, embedding_dim, num_embeddings, commitment_cost, num_var, **kwargs):
        self.embedding_dim = int(embedding_dim)
        self.num_embeddings = int(num_embeddings)
        self.commitment_cost = float(commitment_cost)
        self.num_var = num_var
        self.name = kwargs.get("name", "vector_quantizer")
        self.seed = kwargs.get("seed", 0)
        self.built = False
        self.losses: List[float] = []
        self._model = None
        self._e = None            # DeviceArray [V,K,D] code-major
        self.last_indices = None  # DeviceArray [G,B] int32 of the last call

    def _bind(self, model):
        self._model, self.built = model, True

    # ---- reference-layout attributes -----------------------------------------
    def _get(self, name, dev):
        if self._model is not None:
            return self._model._get_tensor(name)
        a = dev.numpy()
        return np.ascontiguousarray(a.transpose(0, 2, 1)) if a.ndim == 3 else a

    def _set(self, name, attr, value):
        value = np.ascontiguousarray(value, np.float32)
        if self._model is not None:
            self._model._set_tensor(name, value)
            return
        v = np.ascontiguousarray(value.transpose(0, 2, 1)) if value.ndim == 3 else value
        cur = getattr(self, attr)
        if cur is None or cur.shape != v.shape:
            setattr(self, attr, _ffi.DeviceArray.from_numpy(_ffi.get_context(), v))
        else:
            cur.upload(v)

    @property
    def embeddings(self) -> np.ndarray:
        return self._get("vq.embeddings", self._e)

    @embeddings.setter
    def embeddings(self, value):
        self._set("vq.embeddings", "_e", value)
        self.built = True

    def build(self, input_shape):
        if len(input_shape) != 3:
            raise ValueError("The input tensor must be rank of 3")
        num_var = int(input_shape[0])
        rng = np.random.default_rng(self.seed)
        # init.VarianceScaling(distribution='uniform') on shape [V, D, K] (core/quantizer.py:35-36)
        e = initialize("variance_scaling_uniform", (num_var, self.embedding_dim, self.num_embeddings), rng)
        self.embeddings = e
        self.built = True

    # ---- shared call machinery -----------------------------------------------
    def _prepare(self, inputs, fts):
        ctx = _ffi.get_context()
        if self._model is not None and self._e is None:
            self._sync_from_model()
        x = inputs if _ffi.is_device_object(inputs) else _ffi.DeviceArray.from_numpy(ctx, _ffi.as_host_f32(inputs))
        zptr, zshape, keep = _ffi.device_pointer(x)
        if len(zshape) != 3:
            raise ValueError("The input tensor must be rank of 3")
        if not self.built:
            self.build(zshape)
        G, B, D = zshape
        V, K, De = self._e.shape
        if D != De:
            raise ValueError(f"final input dimension {D} must equal embedding_dim {De}")
        if fts is None:
            if G != V:
                raise ValueError(f"inputs carry {G} variables, codebook has {V}")
            e = self._e
        else:
            fts = np.asarray(fts, dtype=np.int64).reshape(-1)
            if len(fts) != G:
                raise ValueError("len(fts) must equal the leading input dimension")
            e = _ffi.DeviceArray.from_numpy(ctx, self._e.numpy()[fts])     # tf.gather(w, fts, axis=0)
        return ctx, x, zptr, keep, G, B, D, K, e

    def _assign(self, ctx, zptr, e, G, B, D, K):
        idx = _ffi.DeviceArray(ctx, (G, B), np.int32, zero=False)
        _ffi.check(_ffi.lib().pgmvae_vq_assign(ctx.h, None, zptr, B * D, D, e.ptr, K * D, D, idx.ptr, B,
                                               None, None, G, B, D, K))
        self.last_indices = idx
        return idx

    def _code_only_output(self, idx, K, fts):
        ind = idx.numpy().astype(np.int64)
        if fts is not None:
            return ind                                               # encoding_indices (:56, :159)
        return np.eye(K, dtype=np.float32)[ind]                      # tf.one_hot -> [V,B,K]

    def _quantize(self, ctx, zptr, e, idx, G, B, D, K):
        q = _ffi.DeviceArray(ctx, (G, B, D), np.float32, zero=False)
        st = _ffi.DeviceArray(ctx, (G, B, D), np.float32, zero=False)
        acc = _ffi.DeviceArray(ctx, (1,), np.float64)
        _ffi.check(_ffi.lib().pgmvae_vq_quantize(ctx.h, None, zptr, B * D, D, e.ptr, K * D, D, idx.ptr, B,
                                                 q.ptr, st.ptr, B * D, D, acc.ptr, G, B, D, K))
        e_latent = float(acc.numpy()[0]) / (G * B * D)               # reduce_mean((q - z)^2)
        return q, st, e_latent

    def get_config(self):
        return {"_embedding_dim": self.embedding_dim, "_num_embeddings": self.num_embeddings,
                "_commitment_cost": self.commitment_cost}


class VectorQuantizer(_VQBase):
    """Gradient-trained VQ layer (reference core/quantizer.py:13-71).

    loss = q_latent_loss + commitment_cost * e_latent_loss; output = straight-through."""

    def _sync_from_model(self):
        self._e = _ffi.DeviceArray.from_numpy(
            _ffi.get_context(), np.ascontiguousarray(self._model._get_tensor("vq.embeddings").transpose(0, 2, 1)))

    def call(self, inputs, training=None, code_only=False, fts=None):
        if self._model is not None:
            self._sync_from_model()
        ctx, x, zptr, keep, G, B, D, K, e = self._prepare(inputs, fts)
        idx = self._assign(ctx, zptr, e, G, B, D, K)
        if code_only:
            self.losses = [0.0]
            return self._code_only_output(idx, K, fts)
        q, st, e_latent = self._quantize(ctx, zptr, e, idx, G, B, D, K)
        # q_latent_loss and e_latent_loss are numerically the same mean (core/quantizer.py:50-52)
        self.losses = [e_latent + self.commitment_cost * e_latent]
        self.last_quantized = q
        return st

    __call__ = call


class VectorQuantizerEMA(_VQBase):
    """EMA-updated VQ layer (reference core/quantizer.py:74-176)."""
    ema = True

    def __init__(self, embedding_dim, num_embeddings, commitment_cost, decay, num_var, epsilon=1e-5, **kwargs):
        super().__init__(embedding_dim, num_embeddings, commitment_cost, num_var, **kwargs)
        self.decay = float(decay)
        self.epsilon = float(epsilon)
        self._ema_c = self._ema_w = self._biased_c = self._biased_w = None
        self._step = 0            # TF's hidden local_step of assign_moving_average

    def build(self, input_shape):
        super().build(input_shape)
        ctx = _ffi.get_context()
        V, K, D = self._e.shape
        self._ema_c = _ffi.DeviceArray(ctx, (V, K))                                   # zeros (:113-114)
        self._ema_w = _ffi.DeviceArray.from_numpy(ctx, self._e.numpy())               # assign(embeddings) (:117)
        self._biased_c = _ffi.DeviceArray(ctx, (V, K))
        self._biased_w = _ffi.DeviceArray(ctx, (V, K, D))
        self._step = 0

    def _sync_from_model(self):
        ctx = _ffi.get_context()
        g = self._model._get_tensor
        t = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
        self._e = _ffi.DeviceArray.from_numpy(ctx, t(g("vq.embeddings")))
        self._ema_w = _ffi.DeviceArray.from_numpy(ctx, t(g("vq.ema_w")))
        self._biased_w = _ffi.DeviceArray.from_numpy(ctx, t(g("vq.biased_w")))
        self._ema_c = _ffi.DeviceArray.from_numpy(ctx, g("vq.ema_cluster_size"))
        self._biased_c = _ffi.DeviceArray.from_numpy(ctx, g("vq.biased_c"))

    @property
    def ema_cluster_size(self) -> np.ndarray:
        return self._get("vq.ema_cluster_size", self._ema_c)

    @ema_cluster_size.setter
    def ema_cluster_size(self, value):
        self._set("vq.ema_cluster_size", "_ema_c", value)

    @property
    def ema_w(self) -> np.ndarray:
        return self._get("vq.ema_w", self._ema_w)

    @ema_w.setter
    def ema_w(self, value):
        self._set("vq.ema_w", "_ema_w", value)

    def call(self, inputs, training=None, code_only=False, fts=None):
        if self._model is not None:
            if training and not code_only:
                raise _ffi.PgmvaeError("a model-bound EMA layer is updated by VqVAE.fit / VqVAE(..., training=True)")
            self._sync_from_model()
        ctx, x, zptr, keep, G, B, D, K, e = self._prepare(inputs, fts)
        idx = self._assign(ctx, zptr, e, G, B, D, K)
        if code_only:
            self.losses = [0.0]
            return self._code_only_output(idx, K, fts)
        q, st, e_latent = self._quantize(ctx, zptr, e, idx, G, B, D, K)
        if training:
            if fts is not None:
                raise ValueError("training with a feature subset is not defined by the reference (w is a gathered copy)")
            L = _ffi.lib()
            counts = _ffi.DeviceArray(ctx, (G, K))
            dw = _ffi.DeviceArray(ctx, (G, K, D))
            _ffi.check(L.pgmvae_ema_stats(ctx.h, None, zptr, B * D, D, idx.ptr, B, counts.ptr, K,
                                          dw.ptr, K * D, D, G, B, D, K))
            self._step += 1
            _ffi.check(L.pgmvae_ema_apply(ctx.h, None, counts.ptr, dw.ptr, self._biased_c.ptr, self._biased_w.ptr,
                                          self._ema_c.ptr, self._ema_w.ptr, self._e.ptr, G, K, D, D,
                                          self.decay, self.epsilon, self._step, 1))
        self.losses = [self.commitment_cost * e_latent]
        self.last_quantized = q
        return st

    __call__ = call

    def get_config(self):
        c = super().get_config()
        c.update({"_decay": self.decay, "_epsilon": self.epsilon})
        return c
