"""Host-side mirror of the reference's synthetic datasets (datasets.py:33-98, 163-279):
same class names, constructor arguments, methods and returned dict keys.  All numbers are
produced on the GPU by libvae_b200 (``ops.dataset_batch`` / ``ops.dataset_score``); keys
are device ``uint32[2]`` tensors updated in place so the fused training kernel can advance
the same chain."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops


class DistributionDataset:
    """datasets.py:33-52."""
    name = None

    def __init__(self, seed, device="cuda"):
        self.device = device
        self.key = ops.key_tensor(ops.prng_key(seed), device)          # datasets.py:35
        self.A = None

    @property
    def is_epochs(self):
        return False

    @property
    def shape(self):
        return (self.ndim,)

    @property
    def dimension(self):                                               # datasets.py:26-28
        return int(np.prod(self.shape))

    def get_key(self):
        """``self.key, key = random.split(self.key)`` (datasets.py:50-52): row 0 carried
        (written back in place), row 1 returned."""
        sp = ops.random_split(self.key)
        self.key.copy_(sp[0])
        return sp[1]

    def data_desc(self) -> _lib.VaeDataDesc:
        return _lib.make_data_desc(self.name, self.dim, self.padding_dim, getattr(self, "intrinsic_dim", 3))

    def _batch(self, size):
        return ops.dataset_batch(self.data_desc(), self.get_key(), self.A, size)

    def get_batch(self, size, return_latents=False):
        samps = self._batch(size)
        return (samps, None) if return_latents else samps

    def _scores(self, batch):
        """Device float32[2]; nothing is read back here (a ``.cpu()`` per model per eval event would
        stall the stream between two persistent launches) -- ``VAEModel.write_stats`` defers it."""
        return ops.dataset_score(self.data_desc(), self.A, batch.contiguous())

    def plot_batch(self, batch, fn=None):
        """PNG output is out of scope; RNG side effects are kept."""

    def save(self, fn):
        pass

    def load(self, fn):
        pass


class SphereDataset(DistributionDataset):
    """datasets.py:55-98."""
    name = "sphere"

    def __init__(self, seed, dimension=3, padding_dimension=0, device="cuda"):
        super().__init__(seed, device)
        self.R = 1
        self.dim = dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension

    def score_batch(self, batch):
        s = self._scores(batch)
        return {"Sphere Error": s[0], "Padding Error": s[1]}


class LinearGaussianDataset(DistributionDataset):
    """datasets.py:163-228."""
    name = "linear_gaussian"

    def __init__(self, seed, dimension=3, intrinsic_dimension=3, padding_dimension=0, var_added=0., device="cuda"):
        super().__init__(seed, device)
        self.dim = dimension
        self.intrinsic_dim = intrinsic_dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension
        self.var_added = var_added
        mat = ops.random_normal(self.get_key(), (dimension, intrinsic_dimension))
        while int(torch.linalg.matrix_rank(mat)) != min(self.dim, self.intrinsic_dim):   # datasets.py:174-178
            mat = ops.random_normal(self.get_key(), (dimension, intrinsic_dimension))
        self.A = mat.contiguous()
        self.transformed_cov = self.A @ self.A.T

    def get_batch(self, size, return_latents=False):
        Y = self._batch(size)
        if self.var_added > 0:                                          # datasets.py:189-192
            noise = ops.random_normal(self.get_key(), (size, self.ndim)) * float(np.sqrt(np.float32(self.var_added)))
            Y = Y + noise
        return (Y, None) if return_latents else Y

    def score_batch(self, batch):
        return {"Squared Norm of padding dimensions": self._scores(batch)[0]}


class GaussianDataset(DistributionDataset):
    """datasets.py:101-160 (not reachable from run.py's --dataset choices; kept for callers that construct it
    directly).  A standard normal in `dimension` coordinates, padding drawn N(0, noise_level I) (or zeros).  It has no
    fused generator: batches are composed from ``ops.random_normal`` draws, so a VAEModel on it trains per-op."""
    name = "gaussian"

    def __init__(self, seed, dimension=3, padding_dimension=0, noise_level=0.01, device="cuda"):
        super().__init__(seed, device)
        self.dim = dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension
        self.noise_level = noise_level

    def data_desc(self):
        raise _lib.VaeB200Error("GaussianDataset has no on-device generator description (per-op batches only)")

    def get_batch(self, size, return_latents=False):
        norm = ops.random_normal(self.get_key(), (size, self.dim))
        if self.noise_level > 0:
            # random.multivariate_normal(key, 0, noise_level I): cholesky(cov) @ normal(key, (size, pad))
            pad = ops.random_normal(self.get_key(), (size, self.padding_dim)) * float(np.sqrt(np.float32(self.noise_level)))
        else:
            pad = torch.zeros((size, self.padding_dim), dtype=torch.float32, device=self.device)
        samps = torch.cat([norm, pad], dim=1)
        return (samps, None) if return_latents else samps

    def score_batch(self, batch):
        padding = batch[:, self.dim:]
        mse = (padding * padding).sum(dim=1).mean()
        w_ht = torch.linalg.eigvalsh(torch.cov(batch.T))
        self.get_batch(batch.shape[0])                       # datasets.py:117 draws a comparison batch from the chain
        return {"Squared Norm of padding dimensions": mse, "ground truth eigenvalue": torch.ones_like(w_ht),
                "learnt eigenvalue": w_ht}


class SigmoidDataset(DistributionDataset):
    """datasets.py:230-279."""
    name = "sigmoid"

    def __init__(self, seed, dimension=3, padding_dimension=0, device="cuda"):
        super().__init__(seed, device)
        self.dim = dimension
        self.padding_dim = padding_dimension
        self.ndim = dimension + padding_dimension + 1
        self.A = ops.random_normal(self.get_key(), (self.dim, 1)).contiguous()   # datasets.py:236-238

    def score_batch(self, batch):
        s = self._scores(batch)
        return {"Squared Norm of Padding Dimensions": s[0], "Squared Norm of Manifold Dimension": s[1]}

    def plot_batch(self, batch, fn=None):
        self.get_batch(batch.shape[0])          # datasets.py:265 draws from the dataset chain
