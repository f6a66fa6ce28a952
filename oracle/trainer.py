"""ORACLE (test infrastructure) -- the training loop / RNG event schedule.
See ``oracle/__init__.py`` for who may import this.

Follows ``model.py:207-222`` (train_distribution), ``model.py:153-168``
(compute_stats), ``model.py:139-146`` (plot_epoch), ``model.py:225-228``
(sample_latent), ``vae.py:56`` (init split), ``vae.py:123-141`` (train_one_batch,
compute_model_stats) and ``vae.py:191-201`` (sample_batch).

Model-side splits consume row 0 and carry row 1; dataset-side ``get_key`` carries
row 0 and returns row 1.
"""
from __future__ import annotations

import numpy as np

from . import jax_random as jr
from . import vae_step as vs

N_PRINT = 5000            # model.py:124
N_PLOT = 50000            # model.py:123
EVAL_BATCH = 1000         # model.py:125-126


class OracleTrainer:
    def __init__(self, desc: vs.VaeDesc, dataset, params, n_print=N_PRINT, n_plot=N_PLOT,
                 eval_batch=EVAL_BATCH, dtype=np.float32):
        self.desc = desc
        self.dataset = dataset
        self.dtype = dtype
        self.n_print, self.n_plot, self.eval_batch = n_print, n_plot, eval_batch
        self.key = jr.PRNGKey(0)                                  # model.py:29
        self.vae_key, self.key = jr.model_chain_next(self.key)    # vae.py:56
        self.params = np.asarray(params, dtype).copy()
        self.m = np.zeros_like(self.params)
        self.v = np.zeros_like(self.params)
        self.step = 0
        self.current_epsilon = np.array([desc.epsilon0], dtype=dtype)   # vae.py:49
        self.vae_losses, self.var_enc, self.var_dec = [], [], []
        self.train_losses = []
        self.stats = {}
        self.latent_keys = []
        self.started = False

    # model.py:225-228
    def sample_latent(self, key, n):
        return jr.normal(key, (n, self.desc.latent_dim + self.desc.data_dim))

    def train_one_batch(self, batch):                             # vae.py:123-130
        L = self.desc.latent_dim
        key, self.key = jr.model_chain_next(self.key)
        self.latent_keys.append(key)
        z = self.sample_latent(key, self.desc.batch)
        self.params, self.m, self.v, self.step, loss = vs.train_step(
            self.desc, self.params, self.m, self.v, self.step, batch, z[:, :L], z[:, L:],
            dtype=self.dtype)
        self.vae_losses.append(loss)
        self.train_losses.append(loss)

    def compute_stats(self):                                      # model.py:153-168
        L = self.desc.latent_dim
        key, self.key = jr.model_chain_next(self.key)
        real, _ = self.dataset.get_batch(self.eval_batch, return_latents=True)
        z = self.sample_latent(key, self.eval_batch)              # vae.py:196
        fake = vs.sample(self.desc, self.params, z[:, :L], z[:, L:], self.current_epsilon,
                         dtype=self.dtype)
        loss, dkl, mse, lv, e, _ = vs.loss_terms(self.desc, self.params, real, z[:, :L], z[:, L:],
                                                 dtype=self.dtype)
        self.vae_losses.append(loss)                              # vae.py:136
        self.var_enc.append(lv)
        self.var_dec.append(e)
        self.current_epsilon = e
        stats = {"VAE Loss": loss, "KL divergence": dkl, "mse": mse}
        stats.update(self.dataset.score_batch(fake))
        for k, val in stats.items():                              # model.py:195-205
            self.stats.setdefault(k, []).append(val)
        return stats

    def plot_epoch(self):                                         # model.py:139-146
        key, self.key = jr.model_chain_next(self.key)
        # the sampled batch only feeds a PNG; the RNG side effects are what matter
        for _ in range(self.dataset.plot_draws()):
            self.dataset.get_key()                                # datasets.py:265

    def begin(self):                                              # model.py:208-211
        _, self.key = jr.model_chain_next(self.key)
        self.real_score = self.dataset.score_batch(self.dataset.get_batch(self.eval_batch))
        self.started = True

    def train(self, num_batches, start=0, stop=None, events=True):
        """Runs steps [start, stop) of an N=num_batches schedule."""
        if not self.started:
            self.begin()
        stop = num_batches if stop is None else stop
        for batchnum in range(start, stop):
            if events and batchnum % self.n_print == 0:
                self.compute_stats()
            if events and (batchnum % self.n_plot == 0 or batchnum == num_batches - 1):
                self.plot_epoch()
            batch = self.dataset.get_batch(self.desc.batch)
            self.train_one_batch(batch)
        return self
