"""ORACLE (test infrastructure) -- NumPy restatement of the three synthetic
generators reachable from ``run.py:46-54``.  See ``oracle/__init__.py``.

Follows ``datasets.py:34-52`` (key chain), ``:55-98`` (SphereDataset),
``:163-228`` (LinearGaussianDataset), ``:230-279`` (SigmoidDataset).

Operation order is fixed (sequential, unfused binary32) so that the CUDA generator
can be bit-exact against this file; the order XLA-CPU would use for these tiny
reductions cannot be observed here (parity unpinned, differences <= a few ulp).
"""
from __future__ import annotations

import numpy as np

from . import jax_random as jr

F32 = np.float32

_EXP_P = [F32(c) for c in (1.9875691500e-4, 1.3981999507e-3, 8.3334519073e-3,
                           4.1665795894e-2, 1.6666665459e-1, 5.0000001201e-1)]
_LOG2E = F32(1.44269504088896341)
_C1 = F32(0.693359375)
_C2 = F32(-2.12194440e-4)


def det_expf(x) -> np.ndarray:
    """Deterministic binary32 exp (Cephes expf, unfused), |x| <= 80 assumed."""
    x = np.asarray(x, dtype=F32)
    n = np.floor(_LOG2E * x + F32(0.5)).astype(F32)
    r = x - n * _C1
    r = r - n * _C2
    z = r * r
    p = _EXP_P[0]
    for c in _EXP_P[1:]:
        p = p * r + c
    y = p * z
    y = y + r
    y = y + F32(1.0)
    scale = ((n.astype(np.int32) + 127) << 23).astype(np.uint32).view(F32)
    return (y * scale).astype(F32)


def det_sigmoid(x) -> np.ndarray:
    """``jax.nn.sigmoid`` = 1 / (1 + exp(-x)) with the deterministic exp."""
    x = np.asarray(x, dtype=F32)
    return (F32(1.0) / (F32(1.0) + det_expf(-x))).astype(F32)


class DistributionDataset:
    """``datasets.py:33-52``."""

    is_epochs = False

    def __init__(self, seed):
        self.key = jr.PRNGKey(seed)

    def get_key(self):
        key, self.key = jr.dataset_chain_next(self.key)
        return key

    @property
    def dimension(self):
        return int(np.prod(self.shape))


class SphereDataset(DistributionDataset):
    """``datasets.py:55-98``."""
    name = "sphere"

    def __init__(self, seed, dimension=3, padding_dimension=0):
        super().__init__(seed)
        self.dim = dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension
        self.A = None

    @property
    def shape(self):
        return (self.ndim,)

    def transform(self, g):
        acc = g[:, 0] * g[:, 0]
        for j in range(1, self.dim):
            acc = acc + g[:, j] * g[:, j]
        rad = np.sqrt(acc).astype(F32)
        out = np.zeros((g.shape[0], self.ndim), dtype=F32)
        out[:, :self.dim] = g / rad[:, None]
        return out

    def get_batch(self, size, return_latents=False):
        g = jr.normal(self.get_key(), (size, self.dim))
        samps = self.transform(g)
        return (samps, None) if return_latents else samps

    def score_batch(self, batch):
        real = batch[:, :self.dim].astype(F32)
        padding = batch[:, self.dim:].astype(F32)
        score = (np.linalg.norm(real, axis=1) - 1) ** 2
        pad = np.linalg.norm(padding, axis=1) ** 2
        return {"Sphere Error": F32(score.mean()), "Padding Error": F32(pad.mean())}

    def plot_draws(self):
        return 0


class LinearGaussianDataset(DistributionDataset):
    """``datasets.py:163-228``."""
    name = "linear_gaussian"

    def __init__(self, seed, dimension=3, intrinsic_dimension=3, padding_dimension=0, var_added=0.0):
        super().__init__(seed)
        self.dim = dimension
        self.intrinsic_dim = intrinsic_dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension
        self.var_added = var_added
        mat = jr.normal(self.get_key(), (dimension, intrinsic_dimension))
        while np.linalg.matrix_rank(mat) != min(self.dim, self.intrinsic_dim):
            mat = jr.normal(self.get_key(), (dimension, intrinsic_dimension))
        self.A = mat

    @property
    def shape(self):
        return (self.ndim,)

    def transform(self, X):
        out = np.zeros((X.shape[0], self.ndim), dtype=F32)
        for i in range(self.dim):
            acc = self.A[i, 0] * X[:, 0]
            for k in range(1, self.intrinsic_dim):
                acc = acc + self.A[i, k] * X[:, k]
            out[:, i] = acc
        return out

    def get_batch(self, size, return_latents=False):
        X = jr.normal(self.get_key(), (size, self.intrinsic_dim))
        Y = self.transform(X)
        if self.var_added > 0:
            noise = jr.normal(self.get_key(), (size, self.ndim)) * np.sqrt(F32(self.var_added))
            Y = (Y + noise).astype(F32)
        return (Y, None) if return_latents else Y

    def score_batch(self, batch):
        padding = batch[:, self.dim:].astype(F32)
        return {"Squared Norm of padding dimensions": F32(np.mean(np.sum(np.square(padding), axis=1)))}

    def plot_draws(self):
        return 0


class SigmoidDataset(DistributionDataset):
    """``datasets.py:230-279``."""
    name = "sigmoid"

    def __init__(self, seed, dimension=3, padding_dimension=0):
        super().__init__(seed)
        self.dim = dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension + 1
        self.A = jr.normal(self.get_key(), (self.dim, 1))

    @property
    def shape(self):
        return (self.ndim,)

    def transform(self, z):
        acc = z[:, 0] * self.A[0, 0]
        for j in range(1, self.dim):
            acc = acc + z[:, j] * self.A[j, 0]
        out = np.zeros((z.shape[0], self.ndim), dtype=F32)
        out[:, :self.dim] = z
        out[:, self.dim] = det_sigmoid(acc)
        return out

    def get_batch(self, size, return_latents=False):
        z = jr.normal(self.get_key(), (size, self.dim))
        Y = self.transform(z)
        return (Y, None) if return_latents else Y

    def score_batch(self, batch):
        batch = batch.astype(F32)
        codomain_hat = batch[:, self.dim]                       # (n,)
        codomain = batch[:, :self.dim] @ self.A                 # (n, 1)
        # datasets.py:256-258 broadcasts (n,) against (n,1) -> (n,n); kept as is.
        manifold_error = np.mean(np.square(codomain_hat - codomain))
        padding = batch[:, self.dim + 1:]
        mse = np.mean(np.sum(np.square(padding), axis=1))
        return {"Squared Norm of Padding Dimensions": F32(mse),
                "Squared Norm of Manifold Dimension": F32(manifold_error)}

    def plot_draws(self):
        """``plot_batch`` draws a fresh batch from the dataset chain (``:265``)."""
        return 1


def make_dataset(name, seed, dimension, padding_dimension, intrinsic_dimension=3, var_added=0.0):
    """``run.py:46-54``."""
    if name == "sphere":
        return SphereDataset(seed, dimension, padding_dimension)
    if name == "linear_gaussian":
        return LinearGaussianDataset(seed, dimension, intrinsic_dimension, padding_dimension, var_added)
    if name == "sigmoid":
        return SigmoidDataset(seed, dimension, padding_dimension)
    raise ValueError(name)
