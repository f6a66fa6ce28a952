"""refshim jax.numpy: float64 torch CPU tensors behind the handful of names the reference uses."""
import math

import numpy as _np
import torch

pi = math.pi
float32 = torch.float32
DT = torch.float64


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(_np.asarray(x), dtype=DT)


def array(x, dtype=None):
    return _t(x)


asarray = array


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def sqrt(x):
    return torch.sqrt(_t(x))


def square(x):
    x = _t(x)
    return x * x


def maximum(a, b):
    a = _t(a)
    return torch.maximum(a, _t(b).to(a.dtype))


def minimum(a, b):
    a = _t(a)
    return torch.minimum(a, _t(b).to(a.dtype))


def sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def zeros(shape, dtype=None):
    return torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=DT)


def ones(shape, dtype=None):
    return torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=DT)


def eye(n):
    return torch.eye(n, dtype=DT)


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)


def dot(a, b):
    return _t(a) @ _t(b)


def isfinite(x):
    return torch.isfinite(_t(x))


class _Linalg:
    @staticmethod
    def norm(x, axis=None, keepdims=False):
        x = _t(x)
        if axis is None:
            return torch.linalg.norm(x)
        return torch.linalg.norm(x, dim=axis, keepdim=keepdims)

    @staticmethod
    def matrix_rank(x):
        return int(torch.linalg.matrix_rank(_t(x)))

    @staticmethod
    def inv(x):
        return torch.linalg.inv(_t(x))

    @staticmethod
    def pinv(x):
        return torch.linalg.pinv(_t(x))


linalg = _Linalg()
