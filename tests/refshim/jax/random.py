"""refshim jax.random = oracle/jax_random.py (legacy threefry2x32 layout).
Keys are uint32[2] NumPy arrays (so `a, b = random.split(key)` unpacks rows); normal() returns the
oracle's float32 values widened to float64 torch tensors."""
import numpy as np
import torch

from oracle import jax_random as _jr


def PRNGKey(seed):
    return _jr.PRNGKey(int(seed))


def split(key, num=2):
    return _jr.split(np.asarray(key, np.uint32), num)


def fold_in(key, data):
    return _jr.fold_in(np.asarray(key, np.uint32), int(data))


def normal(key, shape=(), dtype=None):
    return torch.from_numpy(_jr.normal(np.asarray(key, np.uint32), tuple(shape)).astype(np.float64))


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return torch.from_numpy(_jr.uniform(np.asarray(key, np.uint32), tuple(shape), minval, maxval).astype(np.float64))


def logistic(*a, **k):
    raise NotImplementedError("refshim: random.logistic (model.py:229-234) is unreachable from run.py")


def multivariate_normal(*a, **k):
    raise NotImplementedError("refshim: GaussianDataset (datasets.py:101-160) is unreachable from run.py")
