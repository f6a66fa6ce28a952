"""refshim jax.nn.initializers.  `ones`/`zeros` are exact; `lecun_normal` follows jax's DEFINITION
(truncated normal on [-2, 2] scaled to variance 1/fan_in) but the per-parameter keys are this shim's,
not flax's (unverifiable here): initial parameters are an INPUT recorded in the
golden fixtures, never a parity claim."""
import math

import numpy as np
import torch
from scipy import special as _sp

from .. import random as _random

_TRUNC_STD = 0.87962566103423978


def ones(key, shape, dtype=None):
    return torch.ones(tuple(shape), dtype=torch.float64)


def zeros(key, shape, dtype=None):
    return torch.zeros(tuple(shape), dtype=torch.float64)


def normal(stddev=1e-2):
    def init(key, shape, dtype=None):
        return _random.normal(key, shape) * stddev
    return init


def lecun_normal():
    def init(key, shape, dtype=None):
        u = _random.uniform(key, shape).numpy()                      # [0, 1)
        lo, hi = math.erf(-2 / math.sqrt(2)), math.erf(2 / math.sqrt(2))
        t = math.sqrt(2) * _sp.erfinv(np.clip(lo + (hi - lo) * u, lo + 1e-12, hi - 1e-12))
        std = math.sqrt(1.0 / shape[0]) / _TRUNC_STD
        # legacy flax parameters are float32: round so that every consumer starts from the same values
        return torch.from_numpy((t * std).astype(np.float32).astype(np.float64))
    return init
