"""Packed per-variable dense layer on B200 -- host mirror of the reference ``core/dense.py``.

``FatDense`` keeps the reference constructor and call signature (core/dense.py:46-57, :99):
N independent dense layers in one operator, weight ``kernel [V, in, units]``,
``bias [V, 1, units]``, ``outputs = act(inputs[V,B,in] @ kernel + bias)``.  The TensorFlow
ops behind it (batched matmul, BiasAdd, Selu/Sigmoid) are replaced by one grouped CUDA
kernel with a fused bias/activation epilogue (``pgmvae_dense_fwd``).

A layer is either *standalone* (owns its weights in HBM; used for operator-level work and
tests) or *bound* to a ``core.model.VqVAE`` whose device-resident handle owns all weights in
the library's padded layout; in both cases ``.kernel`` / ``.bias`` read and write numpy
arrays in the reference layout.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from pgmvae import _ffi

_ACTS = {None: None, "linear": None, "selu": "selu", "sigmoid": "sigmoid"}


def _compute_fans(shape: Sequence[int]):
    """Keras ``_compute_fans``: leading axes of a rank>2 shape count as receptive field."""
    if len(shape) < 2:
        return shape[0], shape[0]
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    return shape[-2] * rf, shape[-1] * rf


def initialize(name: str, shape: Sequence[int], rng: np.random.Generator) -> np.ndarray:
    """The initialiser distributions the reference selects by name (core/model.py:20,36,
    core/quantizer.py:36); the random stream is numpy's, TensorFlow's is not reproducible."""
    fan_in, fan_out = _compute_fans(shape)
    if name == "zeros":
        return np.zeros(shape, np.float32)
    if name == "he_uniform":
        lim = math.sqrt(6.0 / fan_in)
    elif name == "glorot_uniform":
        lim = math.sqrt(6.0 / (fan_in + fan_out))
    elif name == "variance_scaling_uniform":
        lim = math.sqrt(3.0 / fan_in)
    else:
        raise ValueError(f"unsupported initializer {name!r}")
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


class FatDense:
    """A N-D layer made of many independent dense layers (reference core/dense.py:15-130).

    Input ``(net_num, batch_size, input_dim)`` -> output ``(net_num, batch_size, units)``.
    """

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, bias_constraint=None, **kwargs):
        if activation not in _ACTS:
            raise ValueError(f"unsupported activation {activation!r} (selu, sigmoid or None)")
        for nm, v in (("kernel_regularizer", kernel_regularizer), ("bias_regularizer", bias_regularizer),
                      ("activity_regularizer", activity_regularizer), ("kernel_constraint", kernel_constraint),
                      ("bias_constraint", bias_constraint)):
            if v is not None:
                raise NotImplementedError(f"{nm} is not used on the reference hot path and is not supported")
        self.units = int(units)
        self.activation = _ACTS[activation]
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.name = kwargs.get("name", "fat_dense")
        self.seed = kwargs.get("seed", 0)
        self.built = False
        self._model = None       # bound mode
        self._index = -1
        self._kernel = None      # standalone mode: DeviceArray [V,in,units]
        self._bias = None

    # ---- bound mode ------------------------------------------------------------
    def _bind(self, model, index: int):
        self._model, self._index, self.built = model, index, True

    # ---- standalone mode -------------------------------------------------------
    def build(self, input_shape):
        if len(input_shape) != 3:
            raise ValueError("The input tensor must be rank of 3")
        num_var, _, last_dim = (int(s) for s in input_shape)
        ctx = _ffi.get_context()
        rng = np.random.default_rng(self.seed)
        self._kernel = _ffi.DeviceArray.from_numpy(
            ctx, initialize(self.kernel_initializer, (num_var, last_dim, self.units), rng))
        self._bias = (_ffi.DeviceArray.from_numpy(ctx, initialize(self.bias_initializer, (num_var, 1, self.units), rng))
                      if self.use_bias else None)
        self.built = True

    @property
    def kernel(self) -> np.ndarray:
        if self._model is not None:
            return self._model._get_tensor(f"fd{self._index}.kernel")
        return self._kernel.numpy()

    @kernel.setter
    def kernel(self, value):
        if self._model is not None:
            self._model._set_tensor(f"fd{self._index}.kernel", value)
        else:
            value = np.ascontiguousarray(value, np.float32)
            if self._kernel is None or self._kernel.shape != value.shape:
                self._kernel = _ffi.DeviceArray.from_numpy(_ffi.get_context(), value)
                self.built = True
            else:
                self._kernel.upload(value)

    @property
    def bias(self) -> Optional[np.ndarray]:
        if self._model is not None:
            return self._model._get_tensor(f"fd{self._index}.bias")
        return None if self._bias is None else self._bias.numpy()

    @bias.setter
    def bias(self, value):
        if self._model is not None:
            self._model._set_tensor(f"fd{self._index}.bias", value)
        else:
            value = np.ascontiguousarray(value, np.float32)
            if self._bias is None or self._bias.shape != value.shape:
                self._bias = _ffi.DeviceArray.from_numpy(_ffi.get_context(), value)
            else:
                self._bias.upload(value)

    def call(self, inputs, fts=None):
        """reference core/dense.py:99-111.  inputs [V,B,in] (numpy or device array);
        returns a DeviceArray [V,B,units] (``np.asarray`` / ``.numpy()`` copies it back)."""
        ctx = _ffi.get_context()
        if self._model is not None:
            kernel = _ffi.DeviceArray.from_numpy(ctx, self.kernel)
            bias = _ffi.DeviceArray.from_numpy(ctx, self.bias) if self.use_bias else None
        else:
            if not self.built:
                self.build(np.shape(inputs) if not isinstance(inputs, _ffi.DeviceArray) else inputs.shape)
            kernel, bias = self._kernel, self._bias
        x = inputs if _ffi.is_device_object(inputs) else _ffi.DeviceArray.from_numpy(ctx, _ffi.as_host_f32(inputs))
        xptr, xshape, _keep = _ffi.device_pointer(x)
        if len(xshape) != 3:
            raise ValueError("The input tensor must be rank of 3")
        G, B, fin = xshape
        V, kin, units = kernel.shape
        if kin != fin:
            raise ValueError(f"input_dim {fin} does not match kernel {kernel.shape}")
        out = _ffi.DeviceArray(ctx, (G, B, units), np.float32)
        act = _ffi.ACT_IDS[self.activation]
        L = _ffi.lib()
        if fts is None:
            if G != V:
                raise ValueError(f"inputs carry {G} nets, layer has {V}")
            _ffi.check(L.pgmvae_dense_fwd(ctx.h, None, xptr, B * fin, fin, kernel.ptr, kin * units, units,
                                          bias.ptr if bias is not None else None, units,
                                          out.ptr, B * units, units, G, B, fin, units, act))
        else:
            # tf.gather(self.kernel, fts, axis=0) (core/dense.py:104-105): one launch per selected net
            fts = np.asarray(fts, dtype=np.int64).reshape(-1)
            if len(fts) != G:
                raise ValueError("len(fts) must equal the leading input dimension")
            for i, v in enumerate(fts):
                _ffi.check(L.pgmvae_dense_fwd(
                    ctx.h, None, xptr + 4 * i * B * fin, 0, fin, kernel.ptr + 4 * int(v) * kin * units, 0, units,
                    (bias.ptr + 4 * int(v) * units) if bias is not None else None, 0,
                    out.ptr + 4 * i * B * units, 0, units, 1, B, fin, units, act))
        return out

    __call__ = call

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.units,)

    def get_config(self):
        return {"units": self.units, "activation": self.activation, "use_bias": self.use_bias,
                "kernel_initializer": self.kernel_initializer, "bias_initializer": self.bias_initializer}
