"""refshim flax.nn (legacy, removed in Flax 0.4): Module / Dense / Model as used at networks.py:26-84
and vae.py:57-60,112.

Semantics restated:
* ``class M(Module): def apply(self, x, **kw)``; inside a parent's ``apply``, ``M(x, name=..., **kw)``
  creates the named submodule and RETURNS ITS OUTPUT;
* ``M.partial(**kw)`` binds keyword arguments; explicit call-time keywords override them;
* ``M.init_by_shape(key, [shape, ...])`` runs ``apply`` once on zero inputs of those shapes in init mode
  and returns ``(outputs, params)``; ``self.param(name, shape, init)`` creates (init mode) or reads the
  parameter; parameter trees are nested dicts keyed by submodule / parameter name;
* ``Model(module, params)`` is callable: ``model(*a, **kw) == module.call(params, *a, **kw)``.
Per-parameter init keys: the parent key folded with the first four bytes of the SHA-1 of the child /
parameter name (recalled from legacy flax, medium confidence -- init values are an input, not a claim).
"""
import hashlib

import torch

from jax import random as _random
from jax.nn import initializers as _init


def _fold_in_str(rng, data):
    d = hashlib.sha1(data.encode("utf-8")).digest()
    return _random.fold_in(rng, int.from_bytes(d[:4], byteorder="big"))


class _Frame:
    def __init__(self, params, rng, init):
        self.params, self.rng, self.init = params, rng, init
        self.counters = {}

    def child(self, name):
        if self.init:
            sub = self.params.setdefault(name, {})
            return _Frame(sub, _fold_in_str(self.rng, name), True)
        if name not in self.params:
            raise KeyError(f"refshim: no parameters for submodule {name!r}")
        return _Frame(self.params[name], None, False)


_stack = []


class _ModuleMeta(type):
    def __call__(cls, *args, name=None, **kwargs):
        if not _stack:
            raise RuntimeError("refshim: a flax.nn.Module can only be applied inside init / Model.__call__")
        parent = _stack[-1]
        if name is None:
            i = parent.counters.get(cls.__name__, 0)
            parent.counters[cls.__name__] = i + 1
            name = f"{cls.__name__}_{i}"
        return cls._run(parent.child(name), args, kwargs)


class Module(metaclass=_ModuleMeta):
    _bound = {}

    @classmethod
    def _run(cls, frame, args, kwargs):
        obj = object.__new__(cls)
        obj._frame = frame
        _stack.append(frame)
        try:
            return obj.apply(*args, **{**cls._bound, **kwargs})
        finally:
            _stack.pop()

    @classmethod
    def partial(cls, *, name=None, **kwargs):
        return _ModuleMeta(cls.__name__, (cls,), {"_bound": {**cls._bound, **kwargs}})

    @classmethod
    def init(cls, rng, *args, **kwargs):
        params = {}
        out = cls._run(_Frame(params, rng, True), args, kwargs)
        return out, params

    @classmethod
    def init_by_shape(cls, rng, input_specs, **kwargs):
        def zeros(spec):
            shape = spec[0] if (len(spec) == 2 and isinstance(spec[0], (tuple, list))) else spec
            return torch.zeros(tuple(int(s) for s in shape), dtype=torch.float64)
        return cls.init(rng, *[zeros(s) for s in input_specs], **kwargs)

    @classmethod
    def call(cls, params, *args, **kwargs):
        return cls._run(_Frame(params, None, False), args, kwargs)

    def param(self, name, shape, initializer):
        fr = self._frame
        if fr.init:
            if name in fr.params:
                raise ValueError(f"refshim: parameter {name!r} defined twice")
            fr.params[name] = initializer(_fold_in_str(fr.rng, name), tuple(shape))
        value = fr.params[name]
        if tuple(value.shape) != tuple(shape):
            raise ValueError(f"refshim: parameter {name!r} has shape {tuple(value.shape)}, expected {tuple(shape)}")
        return value


class Dense(Module):
    """flax.nn.Dense: kernel (in, features), y = x @ kernel + bias; lecun_normal / zeros init."""

    def apply(self, inputs, features, bias=True, kernel_init=None, bias_init=None):
        kernel = self.param("kernel", (inputs.shape[-1], features), kernel_init or _init.lecun_normal())
        y = inputs @ kernel
        if bias:
            y = y + self.param("bias", (features,), bias_init or _init.zeros)
        return y


class BatchNorm(Module):
    def apply(self, x, **kw):
        raise NotImplementedError("refshim: BatchNorm is never enabled on the VAE path (networks.py:40-41)")


def sigmoid(x):
    return 1.0 / (1.0 + torch.exp(-x))


class Model:
    """flax.nn.Model(module, params)."""

    def __init__(self, module, params):
        self.module, self.params = module, params

    def __call__(self, *args, **kwargs):
        return self.module.call(self.params, *args, **kwargs)

    def replace(self, **kw):
        return Model(kw.get("module", self.module), kw.get("params", self.params))
