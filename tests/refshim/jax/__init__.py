"""refshim: the part of the `jax` namespace the reference's VAE path uses (see ../README.md)."""
import contextlib

import torch

from . import numpy  # noqa: F401  (jax.numpy)
from . import random  # noqa: F401
from . import nn  # noqa: F401
from . import lax  # noqa: F401
from . import scipy  # noqa: F401



class _Index:
    """jax.ops.index[...]: the index expression itself."""
    def __getitem__(self, idx):
        return idx


class _Ops:
    """jax.ops.index_update / jax.ops.index (legacy JAX; warm start only, vae.py:62-107): functional x[idx] = y."""
    index = _Index()

    @staticmethod
    def index_update(x, idx, y):
        out = x.clone()
        out[idx] = y
        return out


ops = _Ops()


def jit(fn=None, **_kw):
    """jax.jit: tracing/compilation changes no value -> identity."""
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, *a, **k):
    """Only decorates helpers that the VAE path never calls (networks.py:16, utils.py:68-71)."""
    def unavailable(*args, **kwargs):
        raise NotImplementedError("refshim: jax.vmap is not implemented (not on the VAE path)")
    return unavailable


@contextlib.contextmanager
def disable_jit():
    yield


def value_and_grad(fn):
    """jax.value_and_grad(loss_fn)(model) for a flax.nn.Model argument (networks.py:99): returns the
    loss and a Model whose params hold the gradients.  Reverse mode = torch autograd, float64."""
    import flax

    def wrapped(model):
        leaves = []

        def track(tree):
            out = {}
            for k, v in tree.items():
                if isinstance(v, dict):
                    out[k] = track(v)
                else:
                    t = v.detach().clone().requires_grad_(True)
                    leaves.append(t)
                    out[k] = t
            return out
        tracked = track(model.params)
        loss = fn(flax.nn.Model(model.module, tracked))
        grads = torch.autograd.grad(loss, leaves)
        it = iter(grads)

        def rebuild(tree):
            return {k: rebuild(v) if isinstance(v, dict) else next(it).detach() for k, v in tree.items()}
        return loss.detach(), flax.nn.Model(model.module, rebuild(tracked))
    return wrapped
