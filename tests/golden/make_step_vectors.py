"""Writes tests/golden/step_vectors.npz: regression vectors for the VAE step.

    python tests/golden/make_step_vectors.py

PROVENANCE: produced by the float64 NumPy oracle (oracle/vae_step.py, a restatement of
networks.py:61-101 + legacy flax.optim.Adam), NOT by the reference: jax / flax cannot be
imported in this environment (no wheels, no network), so no reference-generated
fixture can exist.  The oracle's backward is pinned independently by torch autograd in float64
(tests/test_oracle_step.py).  These vectors guard both the oracle and the CUDA path against
drift; inputs are float32, outputs float64.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vae_step as vs  # noqa: E402

CASES = {
    # name: VaeDesc kwargs  (shapes of the three reference scripts, small batch to keep the file small)
    "linear_dd3_pd9_L20": dict(batch=16, data_dim=12, latent_dim=20, epsilon0=-1.0, lr=1e-3),
    "sigmoid_dd3_pd3_L6": dict(batch=16, data_dim=7, latent_dim=6, sig_decoder=True, epsilon0=-3.0, lr=1e-4),
    "sphere_dd3_pd3_L6_h24": dict(batch=16, data_dim=6, latent_dim=6, enc_hidden=[24, 24, 24], dec_hidden=[24, 24, 24],
                                  epsilon0=-3.0, lr=1e-4),
    "fixed_var_L4": dict(batch=10, data_dim=5, latent_dim=4, tunable_decoder_var=False, epsilon0=-2.0, lr=1e-3),
}


def main():
    out = {}
    for name, kw in CASES.items():
        d = vs.VaeDesc(**kw)
        rng = np.random.default_rng(abs(hash(name)) % 2 ** 31 if False else sum(map(ord, name)))
        flat = vs.init_params(d, seed=3)
        flat = (flat + 0.05 * rng.standard_normal(flat.size)).astype(np.float32)
        x = rng.standard_normal((d.batch, d.data_dim)).astype(np.float32)
        z = rng.standard_normal((d.batch, d.latent_dim + d.data_dim)).astype(np.float32)
        L = d.latent_dim
        f64 = flat.astype(np.float64)
        loss, grad = vs.loss_and_grad(d, f64, x, z[:, :L], z[:, L:], dtype=np.float64)
        p1, m1, v1, _, _ = vs.train_step(d, f64, np.zeros(d.n_params), np.zeros(d.n_params), 0, x, z[:, :L], z[:, L:],
                                         dtype=np.float64)
        meta = dict(batch=d.batch, data_dim=d.data_dim, latent_dim=d.latent_dim, enc_hidden=list(d.enc_hidden),
                    dec_hidden=list(d.dec_hidden), sig_decoder=bool(d.sig_decoder), tunable=bool(d.tunable_decoder_var),
                    epsilon0=d.epsilon0, lr=d.lr)
        out[f"{name}/meta"] = np.array(json.dumps(meta))
        out[f"{name}/params"] = flat
        out[f"{name}/x"] = x
        out[f"{name}/z"] = z
        out[f"{name}/loss"] = np.float64(loss)
        out[f"{name}/grad"] = np.asarray(grad, np.float64)
        out[f"{name}/params_after"] = np.asarray(p1, np.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
