"""Initial state of a run: parameter init and the PRNG key chains up to the first training
step.  Key chains, generator matrices and the warm-start perturbations come from the library's device RNG
(bit-identical to jax.random); the random Dense kernels are drawn on the host (init is not on the hot path)
from per-leaf keys derived the way legacy flax derives them.

Parameter init (vae.py:57-60 -> ``VAE.init_by_shape(vae_key, ...)``):
* per-leaf key: legacy ``flax.nn`` folds the SHA-1 of every submodule / parameter NAME into the parent's key
  (``fold_in(key, int.from_bytes(sha1(name)[:4], "big"))``), so the kernel of ``Encoder/FC0`` is drawn with
  ``fold(fold(fold(vae_key, "Encoder"), "FC0"), "kernel")`` -- the names are the reference's own
  (networks.py:36,63-77);
* Dense kernels: ``jax.nn.initializers.lecun_normal`` = ``random.truncated_normal(key, -2, 2, shape) *
  sqrt(1 / fan_in) / 0.8796...``, restated in float32 the way jax computes it (uniform on
  [erf(-sqrt 2), erf(sqrt 2)] from the threefry bits, ``sqrt2 * erf_inv(u)``, clamped to the open interval);
* biases zero, ``epsilon_p`` and ``epsilon`` ones (networks.py:69,71).
What pins it: the reference's own source run on tests/refshim (which restates the same flax rule) produced
the ``init_params`` of tests/golden/ref_*.npz; tests/test_host_logic.py requires this module to reproduce them
(leaf order, names, key folding, distribution).  What cannot be pinned offline: that real flax 0.2 / XLA agree
with the recalled rule and with this erf_inv to the last bit -- the step itself takes parameters as an input.
"""
from __future__ import annotations

import hashlib
import math
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


# ---- host restatement of the few jax.random / flax pieces init needs (uint32 / float32 NumPy; ~2 k words per model)
_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (jax/_src/prng.py threefry2x32; Random123 known answers in tests/test_host_logic.py).
    Works in place on private copies of the counters (a 16 M-element kernel is drawn without temporaries)."""
    ks = (_U32(k0), _U32(k1), _U32(_U32(k0) ^ _U32(k1) ^ _U32(0x1BD11BDA)))
    x0 = np.array(x0, _U32, ndmin=1)
    x1 = np.array(x1, _U32, ndmin=1)
    tmp = np.empty_like(x1)
    x0 += ks[0]
    x1 += ks[1]
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 += x1
            np.left_shift(x1, _U32(r), out=tmp)
            x1 >>= _U32(32 - r)
            x1 |= tmp
            x1 ^= x0
        x0 += ks[(i + 1) % 3]
        x1 += _U32(ks[(i + 2) % 3] + _U32(i + 1))
    return x0, x1


def _threefry_counts(key, counts) -> np.ndarray:
    """jax's array form: counters padded to an even length, first half in lane 0, second half in lane 1."""
    counts = np.asarray(counts, _U32).ravel()
    n = counts.size
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, _U32)])
    h = counts.size // 2
    with np.errstate(over="ignore"):
        y0, y1 = _threefry2x32(key[0], key[1], counts[:h], counts[h:])       # (copies the counters)
    return np.concatenate([y0, y1])[:n]


def fold_in(key, data: int) -> np.ndarray:
    """``jax.random.fold_in(key, data)`` = threefry(key, PRNGKey(data))."""
    return _threefry_counts(key, np.array([0, data & 0xFFFFFFFF], _U32))


def fold_in_str(key, name: str) -> np.ndarray:
    """legacy flax.nn ``_fold_in_str``: the first four bytes of the SHA-1 of the name, big endian."""
    return fold_in(key, int.from_bytes(hashlib.sha1(name.encode("utf-8")).digest()[:4], byteorder="big"))


# Giles' single-precision erf_inv (XLA ErfInv32), highest degree first
_ERFINV_CENTRAL = [np.float32(c) for c in (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
                                           -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)]
_ERFINV_TAIL = [np.float32(c) for c in (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
                                        -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)]


def _horner_f32(coef, t: np.ndarray) -> np.ndarray:
    p = np.full_like(t, coef[0])
    for c in coef[1:]:
        p *= t
        p += c
    return p


def _erf_inv_f32(x: np.ndarray) -> np.ndarray:
    F = np.float32
    x = np.asarray(x, F)
    w = np.multiply(x, x)
    np.negative(w, out=w)
    np.log1p(w, out=w)
    np.negative(w, out=w)                                 # w = -log1p(-x * x)
    p = _horner_f32(_ERFINV_CENTRAL, w - F(2.5))
    tail = np.flatnonzero(w.reshape(-1) >= F(5.0))        # |x| > 0.9966: never reached by the [-2, 2] truncation
    if tail.size:
        p.reshape(-1)[tail] = _horner_f32(_ERFINV_TAIL, np.sqrt(w.reshape(-1)[tail]) - F(3.0))
    p *= x
    return p


def truncated_normal(key, lower: float, upper: float, shape) -> np.ndarray:
    """``jax.random.truncated_normal(key, lower, upper, shape)`` in float32 (jax/_src/random.py ``_truncated_normal``)."""
    F = np.float32
    n = int(np.prod(shape))
    sqrt2 = F(np.sqrt(2.0))
    a = F(math.erf(float(F(lower) / sqrt2)))
    b = F(math.erf(float(F(upper) / sqrt2)))
    bits = _threefry_counts(key, np.arange(n, dtype=_U32))
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(F) - F(1.0)            # [0, 1)
    u = np.maximum(a, (f * F(b - a) + a).astype(F))
    out = (sqrt2 * _erf_inv_f32(u)).astype(F)
    lo = np.nextafter(F(lower), F(np.inf))
    hi = np.nextafter(F(upper), F(-np.inf))
    return np.clip(out, lo, hi).reshape(shape)


def lecun_normal(key, shape) -> np.ndarray:
    """``jax.nn.initializers.lecun_normal()`` = variance_scaling(1.0, "fan_in", "truncated_normal"), float32."""
    std = np.float32(np.sqrt(np.float32(1.0) / np.float32(shape[0]))) / np.float32(_TRUNC_STD)
    return (truncated_normal(key, -2.0, 2.0, shape) * std).astype(np.float32)


def leaf_key(vae_key, name: str) -> np.ndarray:
    """Key of the leaf ``Encoder/FC0/kernel``: the module key folded with every path component in turn."""
    key = np.asarray(vae_key, _U32)
    for part in name.split("/"):
        key = fold_in_str(key, part)
    return key


_INIT_CACHE: dict = {}      # (leaves, vae_key) -> flat parameters; every run of a sweep shares PRNGKey(0) (model.py:29)


def init_params(spec: RunSpec, vae_key: np.ndarray) -> np.ndarray:
    shapes = layer_shapes(spec)
    ck = (tuple(shapes), int(vae_key[0]), int(vae_key[1]))
    hit = _INIT_CACHE.get(ck)
    if hit is not None:
        return hit.copy()
    parts = []
    with np.errstate(over="ignore"):
        for name, shape in shapes:
            n = int(np.prod(shape))
            if name.endswith("kernel"):
                parts.append(lecun_normal(leaf_key(vae_key, name), shape).reshape(-1))
            elif name.endswith("bias"):
                parts.append(np.zeros(n, np.float32))
            else:
                parts.append(np.ones(n, np.float32))
    flat = np.concatenate(parts)
    if flat.size <= (1 << 20):
        if len(_INIT_CACHE) >= 64:
            _INIT_CACHE.clear()
        _INIT_CACHE[ck] = flat.copy()
    return flat


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
