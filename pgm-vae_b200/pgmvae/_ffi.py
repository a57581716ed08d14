"""ctypes binding of libpgmvae.so (include/pgmvae.h).

The host side of the drop-in: ``core.dense`` / ``core.quantizer`` / ``core.model`` call the
CUDA library through these thin wrappers.  Buffers cross the boundary as raw device or host
pointers; numpy carries host data, :class:`DeviceArray` (library-allocated HBM) carries
device data, and foreign GPU tensors are accepted through ``__cuda_array_interface__`` or
DLPack (``__dlpack__``) -- PyTorch is an optional tensor carrier, never a compute path.

There is NO CPU fallback: if the shared library is missing or no B200 is visible every
compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libpgmvae.so")

ACT_NONE, ACT_SELU, ACT_SIGMOID = 0, 1, 2
PREC_FP32, PREC_TF32, PREC_BF16 = 0, 1, 2
ACT_IDS = {None: ACT_NONE, "linear": ACT_NONE, "selu": ACT_SELU, "sigmoid": ACT_SIGMOID}


class PgmvaeError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()

_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_u64, _sz, _cp = C.c_uint64, C.c_size_t, C.c_char_p

# name -> (restype, argtypes); every symbol include/pgmvae.h declares
SIGNATURES = {
    "pgmvae_version": (_i, []),
    "pgmvae_last_error": (_cp, []),
    "pgmvae_device_count": (_i, [C.POINTER(_i)]),
    "pgmvae_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "pgmvae_ctx_destroy": (_i, [_vp]),
    "pgmvae_ctx_sync": (_i, [_vp]),
    "pgmvae_ctx_stream": (_vp, [_vp]),
    "pgmvae_ctx_set_precision": (_i, [_vp, _i]),
    "pgmvae_ctx_get_precision": (_i, [_vp]),
    "pgmvae_ctx_launch_count": (_i64, [_vp]),
    "pgmvae_ctx_profile_begin": (_i, [_vp]),
    "pgmvae_ctx_profile_end": (_i, [_vp, _vp, _sz]),
    "pgmvae_malloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "pgmvae_free": (_i, [_vp, _vp]),
    "pgmvae_malloc_host": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "pgmvae_free_host": (_i, [_vp, _vp]),
    "pgmvae_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "pgmvae_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "pgmvae_memcpy_d2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "pgmvae_memset": (_i, [_vp, _vp, _i, _sz, _vp]),
    "pgmvae_timer_start": (_i, [_vp]),
    "pgmvae_timer_stop_ms": (_i, [_vp, C.POINTER(_f)]),
    "pgmvae_dense_fwd": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _vp, _i64, _i,
                              _i, _i, _i, _i, _i]),
    "pgmvae_dense_fwd_sigmoid_mse": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _vp, _i,
                                          _vp, _i64, _i, _vp, _vp, _i, _i, _i, _i, _i, _f]),
    "pgmvae_dense_dgrad": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _i64, _i, _f,
                                _vp, _i64, _i, _i, _i, _i, _i, _i]),
    "pgmvae_dense_wgrad": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64,
                                _i, _i, _i, _i, _i]),
    "pgmvae_vq_assign": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _vp, _vp, _i, _i, _i, _i]),
    "pgmvae_vq_assign_ema": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i,
                                  _i, _i, _i, _i]),
    "pgmvae_vq_assign_rescored": (_i, [_vp, _i, _i, C.POINTER(_i)]),
    "pgmvae_vq_quantize": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _vp, _vp, _i64, _i, _vp,
                                _i, _i, _i, _i]),
    "pgmvae_vq_codebook_grad": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp, _i64, _vp, _i64, _i, _f, _i, _i, _i, _i]),
    "pgmvae_ema_stats": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i]),
    "pgmvae_ema_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _i, _i]),
    "pgmvae_pll_count": (_i, [_vp, _vp, _vp, _i64, _vp, _i, _i, _vp, _vp, _i, _i, _i]),
    "pgmvae_cpt": (_i, [_vp, _vp, _vp, _vp, _vp, _i64]),
    "pgmvae_pll_reduce": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pgmvae_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f, _d, _d, _d]),
    "pgmvae_y_to_f32": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i]),
    "pgmvae_model_create": (_i, [_vp, C.POINTER(_i), _i, _i, _i, _d, _d, _d, _i, _i, C.POINTER(_vp)]),
    "pgmvae_model_init": (_i, [_vp, _u64]),
    "pgmvae_model_destroy": (_i, [_vp]),
    "pgmvae_model_tensor_size": (_i, [_vp, _cp, C.POINTER(_i64)]),
    "pgmvae_model_set_tensor": (_i, [_vp, _cp, _vp, _i64]),
    "pgmvae_model_get_tensor": (_i, [_vp, _cp, _vp, _i64]),
    "pgmvae_model_set_ema_steps": (_i, [_vp, _i, _i]),
    "pgmvae_model_set_adam_step": (_i, [_vp, _i64]),
    "pgmvae_model_p2p_export": (_i, [_vp, _vp]),
    "pgmvae_model_p2p_import": (_i, [_vp, _i, _i, _vp]),
    "pgmvae_model_p2p_disable": (_i, [_vp]),
    "pgmvae_model_p2p_state_sharded": (_i, [_vp]),
    "pgmvae_p2p_shard_bounds": (_i, [_i, _i, _i, _i, _vp, _vp]),
    "pgmvae_model_p2p_sync_state": (_i, [_vp]),
    "pgmvae_ctx_reserve_sms": (_i, [_vp, _i]),
    "pgmvae_device_can_access_peer": (_i, [_i, _i, _vp]),
    "pgmvae_model_train_step": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _i, _vp]),
    "pgmvae_model_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "pgmvae_model_encode": (_i, [_vp, _vp, _i, _i, _vp]),
    "pgmvae_model_count": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "pgmvae_model_count_vars": (_i, [_vp, _vp, _i, _i64, _i, _i, _vp, _vp]),
    "pgmvae_model_arithmetic": (_i, [_vp]),
    "pgmvae_model_fts_encode": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "pgmvae_model_gibbs_cmll": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, C.c_uint64, _vp, _vp]),
    "pgmvae_model_count_begin": (_i, [_vp]),
    "pgmvae_model_count_add": (_i, [_vp, _vp, _i, _i64, _i, _i]),
    "pgmvae_model_count_end": (_i, [_vp, _i, _i, _vp, _vp]),
    "pgmvae_model_device_bytes": (_i64, [_vp]),
    "pgmvae_model_group_size": (_i, [_vp]),
    "pgmvae_comm_unique_id": (_i, [_vp]),
    "pgmvae_comm_create": (_i, [_vp, _i, _i, _vp, C.POINTER(_vp)]),
    "pgmvae_comm_destroy": (_i, [_vp]),
    "pgmvae_comm_allreduce_f32": (_i, [_vp, _vp, _i64, _vp]),
    "pgmvae_comm_allreduce_f64": (_i, [_vp, _vp, _i64, _vp]),
    "pgmvae_comm_allreduce_u64": (_i, [_vp, _vp, _i64, _vp]),
}


def load_library(path: Optional[str] = None):
    """dlopen libpgmvae.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("PGMVAE_LIB", LIB_PATH)
        if not os.path.exists(p):
            raise PgmvaeError(
                f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
                f"or pgm-vae_b200/build.sh).  There is no CPU fallback.")
        lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def lib():
    return _lib if _lib is not None else load_library()


def check(rc: int):
    if rc != 0:
        msg = lib().pgmvae_last_error()
        raise PgmvaeError(f"libpgmvae error {rc}: {msg.decode() if msg else ''}")


def device_count() -> int:
    n = _i(0)
    check(lib().pgmvae_device_count(C.byref(n)))
    return n.value


# --------------------------------------------------------------------------- context
class Context:
    """pgmvae_ctx: one CUDA device + stream (run.py:27-31 device selection)."""

    def __init__(self, device: int = 0):
        if device < 0:
            raise PgmvaeError("device -1 (the reference's CPU path) is not supported: this build has no CPU fallback")
        self.h = _vp()
        check(lib().pgmvae_ctx_create(int(device), C.byref(self.h)))
        self.device = device
        # PGMVAE_PRECISION = fp32 | tf32 | bf16 selects the arithmetic of the GEMM-shaped kernels for this process
        # (default: the library's exact-fp32 CUDA-core path; run.py asks for tf32 unless told otherwise)
        prec = os.environ.get("PGMVAE_PRECISION", "").lower()
        if prec:
            table = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16, "f16": PREC_BF16}
            if prec not in table:
                raise PgmvaeError(f"PGMVAE_PRECISION={prec!r}: expected fp32, tf32 or bf16")
            self.set_precision(table[prec])

    def sync(self):
        check(lib().pgmvae_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return lib().pgmvae_ctx_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return int(lib().pgmvae_ctx_launch_count(self.h))

    def set_precision(self, prec: int):
        check(lib().pgmvae_ctx_set_precision(self.h, int(prec)))

    def get_precision(self) -> int:
        return int(lib().pgmvae_ctx_get_precision(self.h))

    def reserve_sms(self, n: int):
        """leave n SMs to the communication kernels (data parallel); see pgmvae_ctx_reserve_sms"""
        check(lib().pgmvae_ctx_reserve_sms(self.h, int(n)))

    def profile_begin(self):
        check(lib().pgmvae_ctx_profile_begin(self.h))

    def profile_end(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        check(lib().pgmvae_ctx_profile_end(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def timer_start(self):
        check(lib().pgmvae_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = _f(0)
        check(lib().pgmvae_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            lib().pgmvae_ctx_destroy(self.h)
            self.h = _vp()


_ctxs = {}


def get_context(device: Optional[int] = None) -> Context:
    if device is None:
        device = int(os.environ.get("PGMVAE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _ctxs:
        _ctxs[device] = Context(device)
    return _ctxs[device]


# --------------------------------------------------------------------------- buffers
class DeviceArray:
    """n-d array in HBM allocated through the library (pgmvae_malloc)."""

    def __init__(self, ctx: Context, shape: Sequence[int], dtype=np.float32, zero: bool = True):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = _vp()
        check(lib().pgmvae_malloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value or 0
        self._owner = True
        if zero and self.nbytes:
            check(lib().pgmvae_memset(ctx.h, self.ptr, 0, self.nbytes, None))

    @classmethod
    def from_numpy(cls, ctx: Context, a: np.ndarray) -> "DeviceArray":
        a = np.ascontiguousarray(a)
        d = cls(ctx, a.shape, a.dtype, zero=False)
        d.upload(a)
        return d

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        if self.nbytes:
            check(lib().pgmvae_memcpy_h2d(self.ctx.h, self.ptr, a.ctypes.data, self.nbytes, None))
            self.ctx.sync()          # the pageable source may be released by the caller

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            check(lib().pgmvae_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes, None))
        return out

    def fill_zero(self):
        if self.nbytes:
            check(lib().pgmvae_memset(self.ctx.h, self.ptr, 0, self.nbytes, None))

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3,
                "strides": None, "stream": self.ctx.stream or None}

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __del__(self):
        try:
            if getattr(self, "_owner", False) and self.ptr and self.ctx.h:
                lib().pgmvae_free(self.ctx.h, self.ptr)
        except Exception:
            pass
        self.ptr = 0


# DLPack consumer (no torch import): PyCapsule "dltensor" -> DLManagedTensor
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


_KDLCUDA, _KDLCPU = 2, 1


def _from_dlpack_capsule(cap) -> Tuple[int, Tuple[int, ...], bool]:
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    p = C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
    t = C.cast(p, C.POINTER(_DLTensor)).contents
    shape = tuple(t.shape[i] for i in range(t.ndim))
    if bool(t.strides):
        exp = 1
        for i in reversed(range(t.ndim)):
            if shape[i] != 1 and t.strides[i] != exp:
                raise PgmvaeError("DLPack tensor must be contiguous")
            exp *= shape[i]
    return (t.data or 0) + t.byte_offset, shape, t.device.device_type == _KDLCUDA


def device_pointer(obj) -> Tuple[int, Tuple[int, ...], object]:
    """(device pointer, shape, keep-alive) of a DeviceArray, a __cuda_array_interface__ object
    (torch / cupy tensor) or a DLPack exporter."""
    if isinstance(obj, DeviceArray):
        return obj.ptr, obj.shape, obj
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            raise PgmvaeError("device tensor must be contiguous")
        return int(cai["data"][0]), tuple(cai["shape"]), obj
    if hasattr(obj, "__dlpack__"):
        cap = obj.__dlpack__()
        ptr, shape, is_cuda = _from_dlpack_capsule(cap)
        if not is_cuda:
            raise PgmvaeError("DLPack tensor is not on a CUDA device")
        return ptr, shape, (obj, cap)
    raise PgmvaeError(f"cannot take a device pointer of {type(obj)!r}")


def is_device_object(obj) -> bool:
    if isinstance(obj, DeviceArray) or hasattr(obj, "__cuda_array_interface__"):
        return True
    if isinstance(obj, np.ndarray):
        return False
    dev = getattr(obj, "__dlpack_device__", None)
    if dev is not None:
        try:
            return int(dev()[0]) == _KDLCUDA
        except Exception:
            return False
    return False


def as_host_f32(x) -> np.ndarray:
    if isinstance(x, DeviceArray):
        return x.numpy().astype(np.float32, copy=False)
    if hasattr(x, "detach") and hasattr(x, "cpu"):      # torch tensor carrier
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


def find_nccl() -> Optional[str]:
    """Path of the NCCL shipped with torch (newer than the system one), if any."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return None
