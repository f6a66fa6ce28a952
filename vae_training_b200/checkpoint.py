"""Flat parameter buffer <-> the nested dict ``flax.serialization.to_state_dict(optimizer)``
produces (model.py:85-89; layout per SURVEY.md Appendix D, medium confidence [3P])."""
from __future__ import annotations

import pickle

import numpy as np

from .init import layer_shapes


def _nest(flat_dict):
    out = {}
    for name, val in flat_dict.items():
        node = out
        parts = name.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val
    return out


def unflatten(spec, flat):
    flat = np.asarray(flat)
    out, off = {}, 0
    for name, shape in layer_shapes(spec):
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape).copy()
        off += n
    return out


def flatten(spec, tree):
    def get(name):
        node = tree
        for p in name.split("/"):
            node = node[p]
        return np.asarray(node, np.float32).ravel()
    return np.concatenate([get(name) for name, _ in layer_shapes(spec)])


def to_state_dict(spec, params, adam_m, adam_v, step):
    p, m, v = unflatten(spec, params), unflatten(spec, adam_m), unflatten(spec, adam_v)
    states = {k: {"grad_ema": m[k], "grad_sq_ema": v[k]} for k in p}
    return {"target": {"params": _nest(p)}, "state": {"step": int(step), "param_states": _nest(states)}}


def from_state_dict(spec, sd):
    params = flatten(spec, sd["target"]["params"])
    names = [n for n, _ in layer_shapes(spec)]

    def leaf(name, which):
        node = sd["state"]["param_states"]
        for p in name.split("/"):
            node = node[p]
        return np.asarray(node[which], np.float32).ravel()
    m = np.concatenate([leaf(n, "grad_ema") for n in names])
    v = np.concatenate([leaf(n, "grad_sq_ema") for n in names])
    return params, m, v, int(sd["state"]["step"])


def save_model(fn, spec, params, adam_m, adam_v, step):
    with open(fn, "wb") as f:
        pickle.dump(to_state_dict(spec, params, adam_m, adam_v, step), f)


def load_model(fn, spec):
    with open(fn, "rb") as f:
        return from_state_dict(spec, pickle.load(f))
