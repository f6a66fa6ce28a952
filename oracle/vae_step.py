"""ORACLE (test infrastructure) -- NumPy restatement of the VAE forward, the sampled
ELBO loss, its hand-derived reverse-mode gradient and the legacy flax Adam update.
See ``oracle/__init__.py`` for who may import this.

Follows:
* ``networks.py:26-44``  FullyConnectedNetwork.apply (Dense, relu between layers,
                         optional final sigmoid)
* ``networks.py:59-84``  VAE.apply (encoder -> mu; logvar is the (L,) parameter
                         ``epsilon_p``; tunable decoder log-variance
                         ``epsilon = param(1,) * epsilon0``; reparameterise; one or
                         two decoders; decoder noise)
* ``networks.py:87-101`` VAE.train_step (KL, Gaussian NLL, mean, grad, Adam)
* ``networks.py:103-113`` VAE.loss
* ``vae.py:51-54``       layer size derivation
* ``vae.py:113``         ``flax.optim.Adam(learning_rate)`` [3P: beta1 .9, beta2 .999,
                         eps 1e-8, no weight decay; eps added after sqrt(v_hat)]

Flat parameter layout (shared with include/vae_b200.h): for each of Encoder,
Decoder, [SigDecoder]: for each layer FCi: kernel (in, out) row-major, then bias
(out,); then ``epsilon_p`` (L,), then ``epsilon`` (1,) when tunable.

``dtype`` selects float32 (the oracle proper) or float64 (shadow, for checking the
hand-written backward against autograd).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))


@dataclass
class VaeDesc:
    batch: int
    data_dim: int
    latent_dim: int
    enc_hidden: List[int] = field(default_factory=list)
    dec_hidden: List[int] = field(default_factory=list)
    sig_decoder: bool = False            # dataset_name == "sigmoid" (networks.py:75-78)
    tunable_decoder_var: bool = True     # -tdv
    epsilon0: float = 0.0                # --epsilon
    lr: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8

    @property
    def enc_sizes(self):                 # vae.py:53
        return [self.data_dim] + list(self.enc_hidden) + [self.latent_dim]

    @property
    def dec_sizes(self):                 # vae.py:54
        return [self.latent_dim] + list(self.dec_hidden) + [self.data_dim]

    def segments(self):
        """[(name, shape, offset)] in flat-layout order."""
        segs = []
        off = 0
        nets = [("Encoder", self.enc_sizes), ("Decoder", self.dec_sizes)]
        if self.sig_decoder:
            nets.append(("SigDecoder", self.dec_sizes))
        for net, sizes in nets:
            for i in range(len(sizes) - 1):
                segs.append((f"{net}/FC{i}/kernel", (sizes[i], sizes[i + 1]), off))
                off += sizes[i] * sizes[i + 1]
                segs.append((f"{net}/FC{i}/bias", (sizes[i + 1],), off))
                off += sizes[i + 1]
        segs.append(("epsilon_p", (self.latent_dim,), off))
        off += self.latent_dim
        if self.tunable_decoder_var:
            segs.append(("epsilon", (1,), off))
            off += 1
        return segs

    @property
    def n_params(self):
        name, shape, off = self.segments()[-1]
        return off + int(np.prod(shape))

    def flops_per_step(self):
        """per Dense fwd 2Bmn + dW 2Bmn + dX 2Bmn, no dX for the
        first encoder layer."""
        B = self.batch
        total = 0
        es = self.enc_sizes
        for i in range(len(es) - 1):
            total += (4 if i == 0 else 6) * B * es[i] * es[i + 1]
        ds = self.dec_sizes
        for i in range(len(ds) - 1):
            total += 6 * B * ds[i] * ds[i + 1] * (2 if self.sig_decoder else 1)
        return total


def unflatten(desc: VaeDesc, flat):
    return {name: flat[off:off + int(np.prod(shape))].reshape(shape)
            for name, shape, off in desc.segments()}


def flatten(desc: VaeDesc, tree, dtype=np.float32):
    out = np.zeros(desc.n_params, dtype=dtype)
    for name, shape, off in desc.segments():
        out[off:off + int(np.prod(shape))] = np.asarray(tree[name], dtype=dtype).ravel()
    return out


def init_params(desc: VaeDesc, seed=0, dtype=np.float32):
    """Lecun-normal kernels (truncated +-2 sigma, var 1/fan_in), zero biases, ones for
    ``epsilon_p`` / ``epsilon`` (``vae.py:57-60``, ``networks.py:69,71``).  The values
    follow the *distribution* flax uses, not its exact per-leaf keys (SURVEY.md
    Appendix D) -- initial parameters are an INPUT to the step under test."""
    rng = np.random.default_rng(seed)
    flat = np.zeros(desc.n_params, dtype=dtype)
    for name, shape, off in desc.segments():
        n = int(np.prod(shape))
        if name.endswith("kernel"):
            std = np.sqrt(1.0 / shape[0]) / 0.87962566103423978
            vals = rng.standard_normal(n * 2)
            vals = vals[np.abs(vals) <= 2.0][:n]
            while vals.size < n:
                more = rng.standard_normal(n)
                vals = np.concatenate([vals, more[np.abs(more) <= 2.0]])[:n]
            flat[off:off + n] = (vals * std).astype(dtype)
        elif name.endswith("bias"):
            flat[off:off + n] = 0
        else:
            flat[off:off + n] = 1
    return flat


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _net_names(desc):
    nets = [("Encoder", desc.enc_sizes), ("Decoder", desc.dec_sizes)]
    if desc.sig_decoder:
        nets.append(("SigDecoder", desc.dec_sizes))
    return nets


def _mlp_forward(p, net, sizes, x):
    """Returns (output pre-final-activation, list of layer inputs, list of relu masks)."""
    hs, masks = [], []
    h = x
    nl = len(sizes) - 1
    for i in range(nl):
        hs.append(h)
        h = h @ p[f"{net}/FC{i}/kernel"] + p[f"{net}/FC{i}/bias"]
        if i + 1 < nl:
            masks.append(h > 0)
            h = np.maximum(h, 0)               # utils.py:29-30
    return h, hs, masks


def _mlp_backward(p, net, sizes, hs, masks, dout, grads, need_dx=True):
    nl = len(sizes) - 1
    d = dout
    for i in reversed(range(nl)):
        grads[f"{net}/FC{i}/kernel"] = hs[i].T @ d
        grads[f"{net}/FC{i}/bias"] = d.sum(axis=0)
        if i > 0 or need_dx:
            d = d @ p[f"{net}/FC{i}/kernel"].T
            if i > 0:
                d = d * masks[i - 1]
    return d


def forward(desc: VaeDesc, flat, x, z1, z2, dtype=np.float32, sampling=False, sampling_epsilon=None):
    """``VAE.apply`` (networks.py:61-84).  Returns a dict of intermediates."""
    dt = dtype
    p = unflatten(desc, np.asarray(flat, dtype=dt))
    z1 = np.asarray(z1, dt)
    z2 = np.asarray(z2, dt)
    out = {}
    if sampling:                                   # networks.py:62-65
        mu = np.zeros_like(z1)
        lv = np.zeros(desc.latent_dim, dtype=dt)
        e = np.asarray(sampling_epsilon, dtype=dt).reshape(-1)[:1]
        ehs = emasks = None
    else:
        x = np.asarray(x, dt)
        mu, ehs, emasks = _mlp_forward(p, "Encoder", desc.enc_sizes, x)
        lv = p["epsilon_p"]
        if desc.tunable_decoder_var:
            e = p["epsilon"] * dt(desc.epsilon0)   # networks.py:70-71
        else:
            e = np.full((1,), desc.epsilon0, dtype=dt)
    s = np.exp(lv / dt(2))
    zs = mu + s * z1
    d1, dhs, dmasks = _mlp_forward(p, "Decoder", desc.dec_sizes, zs)
    xhat = d1
    if desc.sig_decoder:
        s_pre, shs, smasks = _mlp_forward(p, "SigDecoder", desc.dec_sizes, zs)
        sg = _sigmoid(s_pre)
        xhat = sg + d1                             # networks.py:76-78
        out.update(sg=sg, shs=shs, smasks=smasks)
    sd = np.exp(e / dt(2))
    xhat = xhat + z2 * sd                          # networks.py:81-83
    out.update(p=p, mu=mu, lv=lv, e=e, s=s, zs=zs, xhat=xhat, sd=sd,
               ehs=ehs, emasks=emasks, dhs=dhs, dmasks=dmasks)
    return out


def loss_terms(desc: VaeDesc, flat, x, z1, z2, dtype=np.float32):
    """``VAE.loss`` (networks.py:103-113): (loss, dkl.mean, mse.mean, logvar_e, epsilon)."""
    dt = dtype
    f = forward(desc, flat, x, z1, z2, dtype=dt)
    x = np.asarray(x, dt)
    mu, lv, e, xhat = f["mu"], f["lv"], f["e"], f["xhat"]
    dkl = dt(-0.5) * np.sum(dt(1) + lv - np.exp(lv) - np.square(mu), axis=-1)
    var_d = np.exp(e)
    mse = (dt(0.5) * np.square(xhat - x) / var_d + dt(0.5) * (dt(LOG_2PI) + e)).sum(axis=-1)
    loss = dkl + mse
    return loss.mean(), dkl.mean(), mse.mean(), lv.copy(), e.copy(), f


def loss_and_grad(desc: VaeDesc, flat, x, z1, z2, dtype=np.float32):
    """value_and_grad of the ``loss_fn`` at networks.py:90-99."""
    dt = dtype
    loss, _, _, _, _, f = loss_terms(desc, flat, x, z1, z2, dtype=dt)
    p = f["p"]
    x = np.asarray(x, dt)
    z1 = np.asarray(z1, dt)
    z2 = np.asarray(z2, dt)
    B = dt(x.shape[0])
    mu, lv, e, s, xhat, sd = f["mu"], f["lv"], f["e"], f["s"], f["xhat"], f["sd"]
    var_d = np.exp(e)
    r = xhat - x
    g = {}
    dxhat = r / (var_d * B)
    de = np.sum(dt(-0.5) * r * r / var_d + dt(0.5) + (r / var_d) * z2 * dt(0.5) * sd) / B
    if desc.tunable_decoder_var:
        g["epsilon"] = np.array([de * dt(desc.epsilon0)], dtype=dt).reshape(1)
    dzs = _mlp_backward(p, "Decoder", desc.dec_sizes, f["dhs"], f["dmasks"], dxhat, g)
    if desc.sig_decoder:
        sg = f["sg"]
        dpre = dxhat * sg * (dt(1) - sg)
        dzs = dzs + _mlp_backward(p, "SigDecoder", desc.dec_sizes, f["shs"], f["smasks"], dpre, g)
    g["epsilon_p"] = np.sum(dzs * z1, axis=0) * dt(0.5) * s - dt(0.5) * (dt(1) - np.exp(lv))
    dmu = dzs + mu / B
    _mlp_backward(p, "Encoder", desc.enc_sizes, f["ehs"], f["emasks"], dmu, g, need_dx=False)
    return dt(loss), flatten(desc, g, dtype=dt)


def adam_update(desc: VaeDesc, flat, m, v, step, grad, dtype=np.float32):
    """Legacy ``flax.optim.Adam.apply_param_gradient`` [3P]; ``step`` is the count of
    updates already applied (starts at 0); returns new (flat, m, v, step+1)."""
    dt = dtype
    b1, b2 = dt(desc.beta1), dt(desc.beta2)
    # hyper-parameters are Python floats in the reference: (1 - beta) is evaluated in double
    # and then rounded to the array dtype; beta ** t is a power in the array dtype
    m = b1 * m + dt(1.0 - desc.beta1) * grad
    v = b2 * v + dt(1.0 - desc.beta2) * np.square(grad)
    t = dt(step + 1)
    m_hat = m / (dt(1) - b1 ** t)
    v_hat = v / (dt(1) - b2 ** t)
    denom = np.sqrt(v_hat) + dt(desc.adam_eps)
    flat = flat - dt(desc.lr) * m_hat / denom
    return flat.astype(dt), m.astype(dt), v.astype(dt), step + 1


def train_step(desc: VaeDesc, flat, m, v, step, x, z1, z2, dtype=np.float32):
    """``VAE.train_step`` (networks.py:87-101): returns (flat, m, v, step, loss)."""
    loss, grad = loss_and_grad(desc, flat, x, z1, z2, dtype=dtype)
    flat, m, v, step = adam_update(desc, np.asarray(flat, dtype), np.asarray(m, dtype),
                                   np.asarray(v, dtype), step, grad, dtype=dtype)
    return flat, m, v, step, loss


def sample(desc: VaeDesc, flat, z1, z2, current_epsilon, dtype=np.float32):
    """Sampling-mode forward (``networks.py:62-65`` via ``vae.py:191-201``)."""
    f = forward(desc, flat, None, z1, z2, dtype=dtype, sampling=True,
                sampling_epsilon=current_epsilon)
    return f["xhat"]
