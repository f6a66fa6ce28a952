"""Data-parallel training of ONE large-batch VAE (BASELINE.json configs[4]) over the GPUs of
a node: one process per GPU, `torch.distributed` (NCCL over NVLink) for the single exchange
step -- a sum-allreduce of the flat gradient (+ the loss scalar) -- and replicated Adam.

Sharding rule: the batch is split by ROW PAIRS {b, b + B/2}.  jax.random's
threefry layout produces flat elements i and i + N/2 from one block, i.e. the same column
of rows b and b + B/2, so a rank that owns whole pairs draws exactly its own eps / dataset
rows with no wasted blocks and bit-identical values to the single-GPU run.  The loss is
`mean` over the GLOBAL batch (networks.py:98): each rank scales by 1/B_global
(`grad_scale_rows`), so the allreduce is a plain sum.  Summation order differs from the
single-GPU run => float32-tolerance parity, not bitwise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from ._lib import VaeB200Error
from .packed import RunSpec


def shard_rows(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) within the first half of the batch owned by `rank`; the rank also owns the
    partner rows [batch/2 + lo, batch/2 + hi).  Halves are split as evenly as possible."""
    if batch % 2:
        raise VaeB200Error("data-parallel sharding needs an even batch size")
    if not 0 <= rank < world:
        raise VaeB200Error("rank out of range")
    half = batch // 2
    base, rem = divmod(half, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_row_indices(batch: int, lo: int, hi: int) -> np.ndarray:
    """Global row index of every local row of a shard, in local order (low rows, partners)."""
    return np.concatenate([np.arange(lo, hi), batch // 2 + np.arange(lo, hi)])


def allreduce_sum(buf: torch.Tensor, group=None) -> torch.Tensor:
    """The path's one collective: in-place sum over ranks (NCCL on GPU; gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class DataParallelVAE:
    """One model, global batch `spec.batch_size`, sharded over the process group.  Keys,
    parameters and Adam state are replicated (every rank performs the same splits and the
    same update); only the batch rows differ."""

    def __init__(self, spec: RunSpec, params: np.ndarray, model_key, dataset_key, A=None, group=None,
                 device="cuda", rank: Optional[int] = None, world: Optional[int] = None, overlap: bool = True):
        self.spec = spec
        self.group = group
        self.device = device
        # rank / world may be given explicitly (single-process replica, or emulated ranks in tests)
        self.collective = rank is None and world is None
        if self.collective:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        else:
            self.world, self.rank = int(world), int(rank)
        self.B = spec.batch_size
        self.lo, self.hi = shard_rows(self.B, self.rank, self.world)
        self.rows = 2 * (self.hi - self.lo)
        self.desc = spec.desc()
        self.data = spec.data_desc()
        n = ops.param_count(self.desc)
        if np.asarray(params).shape != (n,):
            raise VaeB200Error(f"params must have shape ({n},)")
        self.n = n
        self.params = torch.from_numpy(np.asarray(params, np.float32).copy()).to(device)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.key = ops.key_tensor(model_key, device)
        self.dkey = ops.key_tensor(dataset_key, device)
        self.A = None if A is None else torch.from_numpy(np.asarray(A, np.float32).copy()).to(device)
        self.gbuf = torch.zeros(n + 1, dtype=torch.float32, device=device)      # [gradient | loss]
        self.ws = ops.Workspace(device)
        # Overlapped exchange: the backward pass records one event per gradient segment (a layer's
        # kernel + bias, in completion order); a side stream allreduces each segment as it lands and runs Adam on
        # it, while the main stream is still in the earlier layers' backward GEMMs.
        self.segments = ops.grad_segments(self.desc)
        self.overlap = overlap and torch.cuda.is_available() and str(device).startswith("cuda")
        if self.overlap:
            self.comm_stream = torch.cuda.Stream(device=device)
            self.events = [torch.cuda.Event() for _ in self.segments]
            for e in self.events:
                e.record()                 # torch creates the cudaEvent lazily: make the handles exist
            self.done = torch.cuda.Event()

    def draw(self):
        """This rank's rows of the step's dataset batch and eps (model.py:221, vae.py:125-126)."""
        sp = ops.random_split(self.dkey)                  # datasets.py:50-52
        self.dkey.copy_(sp[0])
        x = ops.dataset_batch_pairs(self.data, sp[1], self.A, self.B, self.lo, self.hi)
        sm = ops.random_split(self.key)                   # vae.py:125
        self.key.copy_(sm[1])
        z = ops.random_normal_pairs(sm[0], self.B, self.spec.latent_dimension + self.spec.data_dim, self.lo, self.hi)
        return x, z

    def train_step(self) -> torch.Tensor:
        """One global step; returns the global loss (device scalar view, valid until the next step)."""
        if self.overlap and self.rows > 0:
            return self._train_step_overlapped()
        x, z = self.draw()
        grad, loss = self.gbuf[:self.n], self.gbuf[self.n:]
        if self.rows > 0:
            ops.loss_and_grad(self.desc, self.params, x, z, self.rows, self.B, loss, grad, self.ws)
        else:
            self.gbuf.zero_()
        if self.collective:
            allreduce_sum(self.gbuf, self.group)          # the exchange step
        ops.adam_update(self.desc, self.params, self.adam_m, self.adam_v, self.step, grad)
        return loss

    def _train_step_overlapped(self) -> torch.Tensor:
        """The same step with the exchange bucketed per gradient segment and overlapped with the backward pass:
        segment k is allreduced (NCCL, side stream) and Adam-updated as soon as its event fires, while the main
        stream computes the gradients of the layers below it.  Summation order inside a segment is NCCL's, exactly
        as in the single-allreduce form; parameters a later GEMM of this step still reads are never touched early
        (a layer's dX is enqueued before its segment's event)."""
        main = torch.cuda.current_stream()
        x, z = self.draw()
        grad, loss = self.gbuf[:self.n], self.gbuf[self.n:]
        ops.loss_and_grad_events(self.desc, self.params, x, z, self.rows, self.B, loss, grad, self.ws, self.events)
        with torch.cuda.stream(self.comm_stream):
            for (b, e), ev in zip(self.segments, self.events):
                self.comm_stream.wait_event(ev)
                if self.collective:
                    allreduce_sum(self.gbuf[b:e], self.group)
                ops.adam_update_range(self.desc, self.params, self.adam_m, self.adam_v, self.step, grad, b, e)
            if self.collective:
                allreduce_sum(loss, self.group)
            ops.adam_step_increment(self.step)
            self.done.record(self.comm_stream)
        main.wait_event(self.done)                         # the next step reads the updated parameters
        return loss
