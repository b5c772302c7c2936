"""Shared plumbing of the drop-in modules: tensors in / out of the C ABI, host randomness.

Tensors are torch CUDA tensors (device memory + streams are torch's; every kernel is in
libiris.so).  Anything array-like is accepted and moved to the engine's device; results are
torch CUDA tensors, which export DLPack (``torch.utils.dlpack.to_dlpack`` /
``tensor.__dlpack__``) for ``tf.experimental.dlpack.from_dlpack`` on the TensorFlow side.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .engine import get_engine

_rng = np.random.default_rng()


def set_seed(seed):
    """Seed the host generator behind every random draw of the drop-in functions
    (stands in for ``tf.random.set_seed``)."""
    global _rng
    _rng = np.random.default_rng(seed)


def rng():
    return _rng


def dev(x, dtype=None):
    """array-like -> contiguous fp32 CUDA tensor on the engine's device."""
    import torch
    eng = get_engine()
    t = torch.as_tensor(x)
    t = t.to(device=eng.device, dtype=dtype or torch.float32)
    return t.contiguous()


def empty(shape, like=None):
    import torch
    eng = get_engine()
    return torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device=eng.device)


def call(name, *args):
    """``iris_op_*(ctx, *args, stream)`` on the current torch stream."""
    eng = get_engine()
    fn = getattr(eng.lib, name)
    L.check(fn(eng._ctx, *args, eng._stream()))


def ptr(t):
    return C.c_void_p(t.data_ptr())


def i32_host(a):
    a = np.ascontiguousarray(a, np.int32)
    return a, a.ctypes.data_as(C.c_void_p)


def f32_host(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def norm_axis(axis, ndim):
    return axis + ndim if axis < 0 else axis


def split_axis(shape, axis):
    """-> (outer, n_axis, inner) of a C-contiguous tensor viewed around ``axis``."""
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, int(shape[axis]), inner
