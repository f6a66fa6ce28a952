"""refshim flax.serialization (model.py:37-43, 85-89): Optimizer <-> nested dict of NumPy arrays, in the
layout legacy flax used (SURVEY.md Appendix D, medium confidence): {'target': {'params': ...},
'state': {'step': int, 'param_states': {leaf: {'grad_ema', 'grad_sq_ema'}}}}."""
import numpy as np


def _np(tree):
    return {k: _np(v) if isinstance(v, dict) else v.detach().numpy().astype(np.float32) for k, v in tree.items()}


def _pair(m, v):
    return {k: _pair(m[k], v[k]) if isinstance(m[k], dict) else
            {"grad_ema": m[k].detach().numpy().astype(np.float32), "grad_sq_ema": v[k].detach().numpy().astype(np.float32)}
            for k in m}


def to_state_dict(optimizer):
    return {"target": {"params": _np(optimizer.target.params)},
            "state": {"step": int(optimizer.step), "param_states": _pair(optimizer.grad_ema, optimizer.grad_sq_ema)}}


def from_state_dict(optimizer, state):
    raise NotImplementedError("refshim: nothing in the reference ever calls load()/load_model() (model.py:91-94)")
