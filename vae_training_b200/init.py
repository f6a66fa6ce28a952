"""Initial state of a run: parameter init and the PRNG key chains up to the first training
step, computed with the library's own device RNG (no host re-implementation).

Parameter init follows the *distribution* legacy flax uses (vae.py:57-60; SURVEY.md
Appendix D): Dense kernels lecun-normal (normal truncated at +-2 sigma, variance
1/fan_in), biases zero, ``epsilon_p`` and ``epsilon`` ones (networks.py:69,71).  flax's
per-leaf key derivation cannot be verified here, so values are NOT bit-identical to a
flax run; the step itself takes parameters as an input.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops
from .packed import HostState, RunSpec

_TRUNC_STD = 0.87962566103423978   # std of a unit normal truncated to [-2, 2]


def layer_shapes(spec: RunSpec):
    """[(name, shape)] in the flat layout order of include/vae_b200.h."""
    es = [spec.data_dim] + list(spec.encoder_layer_sizes) + [spec.latent_dimension]   # vae.py:53
    ds = [spec.latent_dimension] + list(spec.layer_sizes) + [spec.data_dim]           # vae.py:54
    nets = [("Encoder", es), ("Decoder", ds)]
    if spec.dataset == "sigmoid":
        nets.append(("SigDecoder", ds))                                               # networks.py:75-77
    out = []
    for net, sizes in nets:
        for i in range(len(sizes) - 1):
            out.append((f"{net}/FC{i}/kernel", (sizes[i], sizes[i + 1])))
            out.append((f"{net}/FC{i}/bias", (sizes[i + 1],)))
    out.append(("epsilon_p", (spec.latent_dimension,)))
    if spec.tunable_decoder_var:
        out.append(("epsilon", (1,)))
    return out


def init_params(spec: RunSpec, vae_key: np.ndarray) -> np.ndarray:
    rng = np.random.default_rng([int(vae_key[0]), int(vae_key[1])])
    parts = []
    for name, shape in layer_shapes(spec):
        n = int(np.prod(shape))
        if name.endswith("kernel"):
            vals = np.empty(0)
            while vals.size < n:
                more = rng.standard_normal(2 * n + 16)
                vals = np.concatenate([vals, more[np.abs(more) <= 2.0]])
            parts.append((vals[:n] * (np.sqrt(1.0 / shape[0]) / _TRUNC_STD)).astype(np.float32))
        elif name.endswith("bias"):
            parts.append(np.zeros(n, np.float32))
        else:
            parts.append(np.ones(n, np.float32))
    return np.concatenate(parts)


def dataset_ctor(spec: RunSpec, device="cuda") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Dataset constructor side effects (datasets.py:34-35, 164-181, 231-238): returns the
    carried dataset key and the generator matrix A (device tensors)."""
    key = ops.key_tensor(ops.prng_key(spec.dataset_seed), device)
    A = None
    if spec.dataset == "linear_gaussian":
        dd, did = spec.dataset_dimension, spec.dataset_intrinsic_dimension
        while True:
            sp = ops.random_split(key)                     # datasets.py:50-52: row 0 carried, row 1 returned
            key, sub = sp[0].contiguous(), sp[1].contiguous()
            A = ops.random_normal(sub, (dd, did))
            if int(torch.linalg.matrix_rank(A)) == min(dd, did):      # datasets.py:174-178
                break
    elif spec.dataset == "sigmoid":
        sp = ops.random_split(key)
        key, sub = sp[0].contiguous(), sp[1].contiguous()
        A = ops.random_normal(sub, (spec.dataset_dimension, 1))       # datasets.py:236-237
    elif spec.dataset != "sphere":
        raise ValueError(f"unknown dataset {spec.dataset!r}")
    return key, A


def initial_state(spec: RunSpec, device="cuda", begun: bool = True) -> HostState:
    """State of a run at the top of the hot loop.  ``begun``: the two pre-loop events of
    model.py:208-209 (model-side split, dataset-side 1000-row batch key) have happened."""
    mkey = ops.key_tensor(ops.prng_key(0), device)                    # model.py:29
    sp = ops.random_split(mkey)                                       # vae.py:56
    vae_key, mkey = sp[0].cpu().numpy(), sp[1].contiguous()
    dkey, A = dataset_ctor(spec, device)
    if begun:
        mkey = ops.random_split(mkey)[1].contiguous()                 # model.py:208
        dkey = ops.random_split(dkey)[0].contiguous()                 # model.py:209 (get_key)
    params = init_params(spec, vae_key)
    return HostState(params=params, adam_m=np.zeros_like(params), adam_v=np.zeros_like(params), step=0,
                     model_key=mkey.cpu().numpy(), dataset_key=dkey.cpu().numpy(),
                     A=None if A is None else A.cpu().numpy())
