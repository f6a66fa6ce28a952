"""Initial state of a run: parameter init and the PRNG key chains up to the first training
step.  Key chains, generator matrices and the warm-start perturbations come from the library's device RNG
(bit-identical to jax.random); the random Dense kernels are drawn on the host with NumPy (see below).

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


def apply_warm_start(spec: RunSpec, params: np.ndarray, key: torch.Tensor, A: Optional[torch.Tensor],
                     latent_off_dimension: int = 0) -> np.ndarray:
    """``warm_start=True`` (vae.py:61-103): overwrite the single-Dense encoder / decoder kernels and ``epsilon_p`` of
    freshly initialised parameters with the reference's hand-built starting point plus small normal perturbations.
    Every perturbation is ``random.normal(self.key, shape)`` with the SAME carried model key (the reference does not
    split it), drawn here with the device RNG; ``key`` is not advanced.  Sphere runs are left unchanged, as in the
    reference.  Shapes the reference's own concatenations cannot produce raise ValueError."""
    if spec.dataset not in ("sigmoid", "linear_gaussian"):
        return params
    if spec.layer_sizes or spec.encoder_layer_sizes:
        raise ValueError("warm start sets Encoder/FC0 and Decoder/FC0 as (D, L) / (L, D) kernels: no hidden layers (vae.py:75-82)")
    L, D, dd = spec.latent_dimension, spec.data_dim, spec.dataset_dimension
    out = np.array(params, np.float32, copy=True)
    offs, o = {}, 0
    for name, shape in layer_shapes(spec):
        offs[name] = (o, shape)
        o += int(np.prod(shape))

    def put(name, value):
        off, shape = offs[name]
        value = np.asarray(value, np.float32)
        if value.shape != tuple(shape):
            raise ValueError(f"warm start: {name} would get shape {value.shape}, the model has {tuple(shape)}")
        out[off:off + value.size] = value.reshape(-1)

    def normal(shape):
        return ops.random_normal(key, shape).cpu().numpy()

    if spec.dataset == "sigmoid":                                    # vae.py:62-82
        if L != D:
            raise ValueError("warm start (sigmoid) needs latent_dimension == dataset.dimension (vae.py:63)")
        eye = np.eye(L, dtype=np.float32)
        eye[dd + 1:, dd + 1:] = 0.0
        enc_var = np.zeros(L, np.float32)
        enc_var[dd + 1:] = -3.0
        put("Decoder/FC0/kernel", eye + normal((L, D)) * np.float32(0.1))
        put("SigDecoder/FC0/kernel", normal((L, D)) * np.float32(0.1))
        put("epsilon_p", enc_var + normal((L,)) * np.float32(0.1))
        put("Encoder/FC0/kernel", eye + normal((D, L)) * np.float32(0.1))
    else:                                                            # vae.py:84-103
        if not dd + latent_off_dimension < L:
            raise ValueError("warm start (linear_gaussian) needs dataset.dim + latent_off_dimension < latent_dimension (vae.py:85)")
        Ah = A.cpu().numpy().astype(np.float32)                      # (dim, intrinsic_dim)
        extra = normal((dd, latent_off_dimension))
        dec = np.concatenate([Ah, extra, np.zeros((dd, L - dd - latent_off_dimension), np.float32)], axis=1)
        dec = np.concatenate([dec, np.zeros((D - dd, L), np.float32)], axis=0)
        if dec.shape != (D, L):
            raise ValueError(f"warm start: decoder constant has shape {dec.shape}, not {(D, L)} (needs intrinsic_dim == dim)")
        dec = dec + normal((D, L)) * np.float32(0.01)
        put("Decoder/FC0/kernel", dec.T)
        enc = np.linalg.pinv(Ah).astype(np.float32)                  # (intrinsic_dim, dim)
        enc = np.concatenate([enc, np.zeros((L - dd, dd), np.float32)], axis=0)
        enc = np.concatenate([enc, np.zeros((L, D - dd), np.float32)], axis=1)
        enc = enc + normal((L, D)) * np.float32(0.01)
        put("Encoder/FC0/kernel", enc.T)
        enc_var = np.zeros(L, np.float32)
        enc_var[:Ah.shape[1] + latent_off_dimension] = -3.0
        put("epsilon_p", enc_var + normal((L,)) * np.float32(0.1))
    return out


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
