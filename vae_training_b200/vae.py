"""Host-side mirror of the reference's training driver for the VAE path:
``Model`` / ``GenerativeModel`` (model.py:18-255) and ``VAEModel`` (vae.py:15-209), same
constructor arguments, method names, event schedule, stat names and save payload.

Every number is computed on the GPU by libvae_b200.  Between two eval events the hot loop
(model.py:221-222) runs inside ONE persistent fused launch when the model's shape has a
fused kernel (``vae_pack_prepare`` once, ``vae_pack_train`` per chunk); otherwise each step is the per-op sequence
get_batch -> split -> normal -> vae_train_step, all enqueued without host synchronisation.
"""
from __future__ import annotations

import os
from collections import defaultdict
from copy import deepcopy
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import checkpoint, init, ops
from ._lib import VaeB200Error
from .packed import RunSpec


def _parse_sizes(s) -> List[int]:
    """vae.py:53-54: "200|200|200" -> [200, 200, 200]; "" -> []."""
    if isinstance(s, (list, tuple)):
        return [int(v) for v in s]
    return [int(v) for v in s.split("|")] if s != "" else []


class VAEModel:
    n_plot = 50000            # model.py:123
    n_print = 5000            # model.py:124
    plot_batch_size = 1000    # model.py:125
    print_batch_size = 1000   # model.py:126

    def __init__(self, dirname, num_batches, num_epochs, batch_size, learning_rate, layer_sizes,
                 encoder_layer_sizes, state_dict, data_fn, epsilon, tqdm, dataset, latent_dimension,
                 tunable_decoder_var=False, warm_start=False, dataset_name=None, latent_off_dimension=0,
                 initial_params: Optional[np.ndarray] = None, device="cuda", verbose=True):
        self.dirname = dirname
        self.num_batches = num_batches
        self.num_epochs = num_epochs
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.state_dict = state_dict
        self.data_fn = data_fn
        self.tqdm = tqdm
        self.verbose = verbose
        self.dataset = dataset
        self.device = device
        self.latent_distribution = "gaussian"                     # vae.py:41
        self.latent_dim = latent_dimension if latent_dimension else dataset.dimension   # model.py:128
        self.latent_dimension = latent_dimension
        self.epsilon = epsilon
        self._stats = defaultdict(list)                           # model.py:35
        self._pending_stats = defaultdict(list)                   # device scalars not yet read back
        self.average_log_likelihoods = []
        self.epoch_num = 0
        self.batchnum = 0
        dname = dataset_name or dataset.name
        self.spec = RunSpec(dataset=dname, dataset_dimension=dataset.dim, padding_dim=dataset.padding_dim,
                            latent_dimension=latent_dimension, layer_sizes=_parse_sizes(layer_sizes),
                            encoder_layer_sizes=_parse_sizes(encoder_layer_sizes), batch_size=batch_size,
                            learning_rate=learning_rate, epsilon=epsilon, tunable_decoder_var=tunable_decoder_var,
                            dataset_intrinsic_dimension=getattr(dataset, "intrinsic_dim", 3),
                            num_batches=num_batches)
        if self.spec.data_dim != dataset.dimension:
            raise VaeB200Error("dataset shape and model description disagree")
        self.desc = self.spec.desc()
        # key chain: PRNGKey(0) (model.py:29), vae_key split (vae.py:56)
        self.key = ops.key_tensor(ops.prng_key(0), device)
        sp = ops.random_split(self.key)
        self.vae_key = sp[0].cpu().numpy()
        self.key.copy_(sp[1])
        params = init.init_params(self.spec, self.vae_key) if initial_params is None else np.asarray(initial_params, np.float32)
        if warm_start:                                            # vae.py:61-103 (perturbations drawn with the carried key)
            params = init.apply_warm_start(self.spec, params, self.key, dataset.A, latent_off_dimension)
        n = ops.param_count(self.desc)
        if params.shape != (n,):
            raise VaeB200Error(f"initial_params must have shape ({n},)")
        self.params = torch.from_numpy(params.copy()).to(device)
        self.adam_m = torch.zeros_like(self.params)               # flax.optim.Adam(...).create (vae.py:113)
        self.adam_v = torch.zeros_like(self.params)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.current_epsilon = torch.full((1,), float(epsilon), dtype=torch.float32, device=device)   # vae.py:49
        self.ws = ops.Workspace(device)
        self._chunk_loss = torch.zeros(self.n_print, dtype=torch.float32, device=device)
        self._scratch = ops.fused_scratch(1, device)
        self._job = ops.make_job(self.desc, dataset.data_desc(), self.params, self.adam_m, self.adam_v, self.step,
                                 self.key, dataset.key, dataset.A, self._chunk_loss)
        self.fused_class = ops.fused_supported(self._job)
        if getattr(dataset, "var_added", 0.0) > 0:      # -dn noise (datasets.py:189-192) exists only in the per-op generator
            self.fused_class = 0
        self.fused = self.fused_class > 0
        self._fused_ws = ops.fused_workspace(self._job, device) if self.fused else None
        if self._fused_ws is not None:
            self._job = ops.make_job(self.desc, dataset.data_desc(), self.params, self.adam_m, self.adam_v, self.step,
                                     self.key, dataset.key, dataset.A, self._chunk_loss, workspace=self._fused_ws)
        self._plan = None
        # vae_losses interleaves train and eval losses (vae.py:130,136): kept as device pieces
        self._loss_pieces: List[torch.Tensor] = []
        self.var_enc: List[torch.Tensor] = []
        self.var_dec: List[torch.Tensor] = []
        self.gt_eigen, self.ht_eigen = [], []
        self.correlation_ratios = []
        off_lv = n - latent_dimension - (1 if tunable_decoder_var else 0)
        self._lv = self.params[off_lv:off_lv + latent_dimension]
        self._eps_param = self.params[n - 1:n] if tunable_decoder_var else None

    # ------------------------------------------------------------------ RNG helpers
    def _split_model_key(self) -> torch.Tensor:
        """``sub, self.key = random.split(self.key)``: row 0 consumed, row 1 carried."""
        sp = ops.random_split(self.key)
        self.key.copy_(sp[1])
        return sp[0]

    def sample_latent(self, key, batch_size):                      # model.py:225-228
        return ops.random_normal(key, (batch_size, self.latent_dim + self.dataset.dimension))

    # ------------------------------------------------------------------ hot path
    def train_one_batch(self, batch):                              # vae.py:123-130
        batch = batch.reshape(batch.shape[0], -1).contiguous()
        latent_batch_key = self._split_model_key()
        z = self.sample_latent(latent_batch_key, self.batch_size)
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        ops.train_step(self.desc, self.params, self.adam_m, self.adam_v, self.step, batch, z, loss, self.ws)
        self._loss_pieces.append(loss)

    def train_steps(self, n_steps: int):
        """n consecutive iterations of model.py:221-222."""
        if self.fused:
            done = 0
            while done < n_steps:
                c = min(self.n_print, n_steps - done)
                if self._plan is None:          # validation + job-table upload happen once per model
                    self._plan = ops.pack_prepare([self._job], self._scratch)
                ops.pack_train(self._plan, c)
                self._loss_pieces.append(self._chunk_loss[:c].clone())
                done += c
        else:
            for _ in range(n_steps):
                self.train_one_batch(self.dataset.get_batch(self.batch_size))
        self.batchnum += n_steps

    # ------------------------------------------------------------------ eval path
    def sample_batch(self, key, batch_size, latents=None):         # vae.py:191-201
        z = latents if latents is not None else self.sample_latent(key, batch_size)
        x_hat = ops.vae_sample(self.desc, self.params, z, z.shape[0], self.current_epsilon, self.ws)
        return x_hat, z

    def compute_model_stats(self, real_batch, fake_batch, latents):   # vae.py:132-141
        out = ops.vae_loss(self.desc, self.params, real_batch.contiguous(), latents, latents.shape[0], self.ws)
        self._loss_pieces.append(out[0:1])
        logvar_e = self._lv.clone()
        if self._eps_param is not None:
            epsilon = self._eps_param * float(np.float32(self.epsilon))      # networks.py:70-71
        else:
            epsilon = torch.full((1,), float(self.epsilon), dtype=torch.float32, device=self.device)
        self.var_enc.append(logvar_e)
        self.var_dec.append(epsilon)
        self.current_epsilon = epsilon.contiguous()
        return {"VAE Loss": out[0], "KL divergence": out[1], "mse": out[2]}

    def compute_stats(self):                                       # model.py:153-168
        key = self._split_model_key()
        real_batch, latents = self.dataset.get_batch(self.print_batch_size, return_latents=True)
        if latents is None or latents.shape[-1] != self.latent_dim:
            latents = None
        fake_batch, latents = self.sample_batch(key, self.print_batch_size, latents=latents)
        stats = self.compute_model_stats(real_batch, fake_batch, latents)
        score = self.dataset.score_batch(fake_batch)
        stats.update(score)
        return stats

    @property
    def stats(self):
        """stat name -> list of floats (model.py:35).  Values recorded by a quiet run stay on the
        device until somebody looks: reading them back at every event would stall the stream
        between two persistent launches."""
        for stat, pend in self._pending_stats.items():
            if pend:
                self._stats[stat].extend(torch.stack([p.reshape(()) for p in pend]).cpu().tolist())
                pend.clear()
        return self._stats

    def write_stats(self, stats):                                  # model.py:195-205
        if not self.verbose:
            for stat, val in stats.items():
                if torch.is_tensor(val) and val.is_cuda:
                    self._pending_stats[stat].append(val.detach().clone())
                else:
                    self.stats[stat].append(float(val))
            return
        message = f"Batch | {self.batchnum}"
        for stat, val in stats.items():
            val = float(val)
            self.stats[stat].append(val)
            message = message + f" | {stat} | {val:.3f}"
        print(message, flush=True)

    def plot_epoch(self):                                          # model.py:139-146
        key = self._split_model_key()
        batch = self.sample_batch(key, self.plot_batch_size)[0]
        fn = os.path.join(self.dirname, f"output_{self.batchnum}.png") if self.dirname else None
        self.dataset.plot_batch(batch, fn=fn)

    def plot(self):                                                # model.py:135-137
        pass

    # ------------------------------------------------------------------ loop
    def begin(self):                                               # model.py:208-211
        self._split_model_key()
        eval_batch = self.dataset.get_batch(self.print_batch_size)
        score = self.dataset.score_batch(eval_batch)
        if self.verbose:
            print(f"Score for real data: { {k: float(v) for k, v in score.items()} }", flush=True)
        self.batchnum = 0

    def events_at(self, batchnum: int, save=True):
        """The two event checks at the top of loop iteration `batchnum` (model.py:214-220)."""
        self.batchnum = batchnum
        if batchnum % self.n_print == 0:
            self.write_stats(self.compute_stats())
        if batchnum % self.n_plot == 0 or batchnum == self.num_batches - 1:
            self.plot_epoch()
            if save:
                self.save()

    def next_event(self, batchnum: int) -> int:
        """Smallest event iteration > batchnum, or num_batches."""
        nxt = min((batchnum // self.n_print + 1) * self.n_print, (batchnum // self.n_plot + 1) * self.n_plot)
        last = self.num_batches - 1
        if batchnum < last < nxt:
            nxt = last
        return min(nxt, self.num_batches)

    def train_distribution(self, save=True):                       # model.py:207-222
        self.begin()
        b = 0
        while b < self.num_batches:
            self.events_at(b, save=save)
            nxt = self.next_event(b)
            self.train_steps(nxt - b)
            b = nxt
        self.batchnum = self.num_batches - 1

    def train(self):                                               # model.py:170-174
        self.train_distribution()

    # ------------------------------------------------------------------ persistence
    @property
    def vae_losses(self) -> np.ndarray:
        if not self._loss_pieces:
            return np.zeros(0, np.float32)
        return torch.cat([p.reshape(-1) for p in self._loss_pieces]).cpu().numpy()

    def model_save_data(self, final=False):                        # vae.py:203-209
        data = {"VAE Loss": self.vae_losses,
                "Decoder Variance": np.array([v.cpu().numpy() for v in self.var_dec]),
                "Encoder Variance": np.array([v.cpu().numpy() for v in self.var_enc]),
                "EigenValues": np.zeros((2, 0))}
        if final:
            data["Correlation Ratio"] = np.zeros(0)                # params_and_gradients is always empty
        return data

    def save_model(self):                                          # model.py:85-89
        checkpoint.save_model(os.path.join(self.dirname, "model.pkl"), self.spec, self.params.cpu().numpy(),
                              self.adam_m.cpu().numpy(), self.adam_v.cpu().numpy(), int(self.step))

    def load_model(self):                                          # model.py:37-43
        if self.state_dict is None:
            return
        p, m, v, step = checkpoint.load_model(self.state_dict, self.spec)
        self.params.copy_(torch.from_numpy(p))
        self.adam_m.copy_(torch.from_numpy(m))
        self.adam_v.copy_(torch.from_numpy(v))
        self.step.fill_(step)

    def save(self, final=False):                                   # model.py:246-255
        if not self.dirname:
            return
        data = self.model_save_data(final=final)
        data["Average Log Likelihood"] = np.array(self.average_log_likelihoods)
        stats = deepcopy(dict(self.stats))
        stats.update(data)
        np.savez(os.path.join(self.dirname, "losses"), **stats)
        self.save_model()
        self.dataset.save(os.path.join(self.dirname, "dataset.pk"))


def train_many(models: Sequence[VAEModel], save=True, timings: Optional[dict] = None):
    """Packed sweep: every model follows its own reference schedule, but all models whose
    shape has a fused kernel advance together, one persistent launch per chunk (one CTA per
    model).  Models stay step-aligned; a chunk ends at the nearest event of any model."""
    other = [m for m in models if not m.fused]
    for m in models:
        m.begin()
    # One pack per fused-kernel class (single-CTA models / cluster models).  The two packs are
    # independent serial chains, so they run CONCURRENTLY: one host thread + one CUDA stream per
    # class (the ctypes calls release the GIL; torch's current stream is thread-local).
    def run_pack(fused, stream, errors, barrier):
        import time
        t_begin = time.perf_counter()
        try:
            with torch.cuda.stream(stream):
                scratch = ops.fused_scratch(len(fused), fused[0].device)
                b = 0
                plan, plan_ids = None, None
                first = True
                while True:
                    active = [m for m in fused if b < m.num_batches]
                    if not active:
                        break
                    for m in active:
                        m.events_at(b, save=save)
                    nxt = min(m.next_event(b) for m in active)
                    c = nxt - b
                    ids = tuple(id(m) for m in active)
                    if ids != plan_ids:         # the pack only changes when a run finishes: re-prepare then
                        plan, plan_ids = ops.pack_prepare([m._job for m in active], scratch), ids
                    if first:
                        # CUDA loads kernels lazily, and loading one while ANOTHER stream's long persistent kernel is
                        # running waits for that kernel to finish: the ~30 kernels of this pack's event path would each
                        # stall behind a 0.8 s hidden-layer chunk of the other pack (measured: a 4-run single-Dense pack
                        # took 29 s instead of 4.6 s beside one sphere chain).  So every pack runs its first events and
                        # prepares its plan (which loads its fused kernel) before ANY pack starts a persistent launch.
                        stream.synchronize()
                        barrier.wait()
                        first = False
                    ops.pack_train(plan, c)
                    for m in active:
                        m._loss_pieces.append(m._chunk_loss[:c].clone())
                        m.batchnum = nxt
                    b = nxt
                stream.synchronize()
            if timings is not None:
                timings[f"class{fused[0].fused_class}_pack_seconds"] = time.perf_counter() - t_begin
                timings[f"class{fused[0].fused_class}_pack_runs"] = len(fused)
        except BaseException as e:          # re-raised on the caller's thread
            barrier.abort()
            errors.append(e)

    packs = [[m for m in models if m.fused_class == cls] for cls in (1, 2)]
    packs = [p for p in packs if p]
    if packs:
        import threading
        dev = packs[0][0].device
        torch.cuda.current_stream(dev).synchronize()           # state uploads are complete
        errors, threads = [], []
        barrier = threading.Barrier(len(packs))
        for p_ in packs:
            st = torch.cuda.Stream(device=dev)
            th = threading.Thread(target=run_pack, args=(p_, st, errors, barrier))
            th.start()
            threads.append(th)
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
    for m in other:
        b = 0
        while b < m.num_batches:
            m.events_at(b, save=save)
            nxt = m.next_event(b)
            m.train_steps(nxt - b)
            b = nxt
    for m in models:
        m.batchnum = m.num_batches - 1
