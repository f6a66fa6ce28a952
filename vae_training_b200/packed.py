"""Packed training of many independent VAE runs on one GPU (the sweep path).

A ``PackedRuns`` object owns the device state of M models (parameters, Adam moments,
step counters, carried PRNG keys, generator matrices, loss history) in a handful of flat
device buffers and advances all of them with ONE persistent launch per chunk of steps
(``vae_train_fused``, include/vae_b200.h).  It is the object ``bench.py`` times and the
engine under ``vae.VAEModel.train`` / ``sweep.run_sweep``.

Host <-> device contract of ``train_host`` (the end-to-end call): inputs are HOST arrays
(initial parameters, keys, generator matrices); they are copied to the device from pinned
memory, the fused kernel runs, and per-step losses plus the final state are copied back.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import VaeB200Error


@dataclass
class RunSpec:
    """One model of a sweep: the flags of one ``python run.py ...`` line that shape the
    hot path (run.py:10-38)."""
    dataset: str                     # --dataset
    dataset_dimension: int           # -dd
    padding_dim: int                 # --padding_dim
    latent_dimension: int            # --latent_dim
    layer_sizes: Sequence[int] = ()          # --layer_sizes        (decoder hidden)
    encoder_layer_sizes: Sequence[int] = ()  # --encoder_layer_sizes
    batch_size: int = 100            # --batch_size
    learning_rate: float = 1e-4      # -lr
    epsilon: float = 0.0             # -e
    tunable_decoder_var: bool = False  # -tdv
    dataset_seed: int = 69           # -ds
    dataset_intrinsic_dimension: int = 3  # -did
    num_batches: int = 15000         # --num_batches
    name: str = ""

    @property
    def data_dim(self) -> int:       # datasets.py:61,169,235
        return self.dataset_dimension + self.padding_dim + (1 if self.dataset == "sigmoid" else 0)

    def desc(self) -> _lib.VaeDesc:
        return _lib.make_desc(self.batch_size, self.data_dim, self.latent_dimension, self.encoder_layer_sizes,
                              self.layer_sizes, self.dataset == "sigmoid", self.tunable_decoder_var, self.epsilon,
                              self.learning_rate)

    def data_desc(self) -> _lib.VaeDataDesc:
        return _lib.make_data_desc(self.dataset, self.dataset_dimension, self.padding_dim,
                                   self.dataset_intrinsic_dimension)

    def flops_per_step(self) -> int:
        """per Dense fwd 2Bmn + dW 2Bmn + dX 2Bmn, no dX for the first
        encoder layer."""
        B = self.batch_size
        es = [self.data_dim] + list(self.encoder_layer_sizes) + [self.latent_dimension]
        ds = [self.latent_dimension] + list(self.layer_sizes) + [self.data_dim]
        tot = sum((4 if i == 0 else 6) * B * es[i] * es[i + 1] for i in range(len(es) - 1))
        tot += sum(6 * B * ds[i] * ds[i + 1] for i in range(len(ds) - 1)) * (2 if self.dataset == "sigmoid" else 1)
        return tot

    def normals_per_step(self) -> int:
        g = self.dataset_intrinsic_dimension if self.dataset == "linear_gaussian" else self.dataset_dimension
        return self.batch_size * (self.latent_dimension + self.data_dim + g)


@dataclass
class HostState:
    """Host-side state of one model (what ``train_host`` takes and returns)."""
    params: np.ndarray
    adam_m: np.ndarray
    adam_v: np.ndarray
    step: int
    model_key: np.ndarray            # uint32[2] carried model-side key
    dataset_key: np.ndarray          # uint32[2] carried dataset-side key
    A: Optional[np.ndarray]          # generator matrix (None for sphere)
    losses: List[np.ndarray] = field(default_factory=list)


def _offsets(sizes, align=4):
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += (int(s) + align - 1) // align * align
    return offs, o


class PackedRuns:
    def __init__(self, specs: Sequence[RunSpec], max_chunk: int, device="cuda"):
        if not specs:
            raise VaeB200Error("PackedRuns needs at least one run")
        self.specs = list(specs)
        self.M = len(self.specs)
        self.device = torch.device(device)
        self.max_chunk = int(max_chunk)
        self.descs = [s.desc() for s in self.specs]
        self.datas = [s.data_desc() for s in self.specs]
        self.n_params = [ops.param_count(d) for d in self.descs]
        self.p_off, ptot = _offsets(self.n_params)
        a_sizes = [self._a_size(s) for s in self.specs]
        self.a_sizes = a_sizes
        self.a_off, atot = _offsets(a_sizes)
        dev = self.device
        # ONE device arena and ONE pinned host mirror, laid out [P | M | V | A | step | keys | loss] (every
        # segment 16-byte aligned, all 4-byte element types), so the end-to-end path moves the whole state
        # with a single H2D copy and state + losses with a single D2H copy.
        na = max(atot, 4)
        seg = [("P", ptot), ("Mo", ptot), ("V", ptot), ("A", na), ("step", self.M * 4), ("keys", self.M * 8),
               ("loss", self.M * self.max_chunk)]
        offs, words = _offsets([n for _, n in seg])
        self._seg = {name: (o, n) for (name, n), o in zip(seg, offs)}
        self._state_words = self._seg["loss"][0]                       # everything before the loss history
        self.arena = torch.zeros(words, dtype=torch.float32, device=dev)
        pin = torch.cuda.is_available()
        self.harena = torch.zeros(words, dtype=torch.float32, pin_memory=pin)

        def view(buf, name, dtype):
            o, n = self._seg[name]
            return buf[o:o + n].view(dtype)
        self.P, self.Mo, self.V, self.A = (view(self.arena, k, torch.float32) for k in ("P", "Mo", "V", "A"))
        self.step = view(self.arena, "step", torch.int32)              # stride 4: 16-byte slots
        self.keys = view(self.arena, "keys", torch.uint32)             # [model_key(2) pad(2) dataset_key(2) pad(2)]
        self.loss = view(self.arena, "loss", torch.float32)
        self.scratch = ops.fused_scratch(self.M, dev)
        self.hP, self.hM, self.hV, self.hA = (view(self.harena, k, torch.float32) for k in ("P", "Mo", "V", "A"))
        self.hstep = view(self.harena, "step", torch.int32)
        self.hkeys = view(self.harena, "keys", torch.uint32)
        self.hloss = view(self.harena, "loss", torch.float32)
        self.jobs = [self._make_job(i) for i in range(self.M)]
        # class-2 (cluster kernel) jobs need an L2-resident exchange workspace each
        self.workspaces = [ops.fused_workspace(j, dev) for j in self.jobs]
        self.jobs = [self._make_job(i, self.workspaces[i]) for i in range(self.M)]
        self.plan = None
        unsupported = [i for i, j in enumerate(self.jobs) if ops.fused_supported(j) == 0]
        if unsupported:
            s = self.specs[unsupported[0]]
            raise VaeB200Error(f"run {unsupported[0]} ({s.dataset} dd={s.dataset_dimension} pd={s.padding_dim} "
                               f"L={s.latent_dimension} hidden={list(s.layer_sizes)}) has no fused kernel")

    @staticmethod
    def _a_size(s: RunSpec) -> int:
        if s.dataset == "linear_gaussian":
            return s.dataset_dimension * s.dataset_intrinsic_dimension
        if s.dataset == "sigmoid":
            return s.dataset_dimension
        return 0

    # views -----------------------------------------------------------------
    def params(self, i):
        return self.P[self.p_off[i]:self.p_off[i] + self.n_params[i]]

    def adam_m(self, i):
        return self.Mo[self.p_off[i]:self.p_off[i] + self.n_params[i]]

    def adam_v(self, i):
        return self.V[self.p_off[i]:self.p_off[i] + self.n_params[i]]

    def gen_matrix(self, i):
        return self.A[self.a_off[i]:self.a_off[i] + self.a_sizes[i]] if self.a_sizes[i] else None

    def step_of(self, i):
        return self.step[4 * i:4 * i + 1]

    def model_key(self, i):
        return self.keys[8 * i:8 * i + 2]

    def dataset_key(self, i):
        return self.keys[8 * i + 4:8 * i + 6]

    def losses(self, i, n):
        return self.loss[i * self.max_chunk:i * self.max_chunk + n]

    def _make_job(self, i, workspace=None):
        return ops.make_job(self.descs[i], self.datas[i], self.params(i), self.adam_m(i), self.adam_v(i),
                            self.step_of(i), self.model_key(i), self.dataset_key(i), self.gen_matrix(i),
                            self.losses(i, self.max_chunk), workspace=workspace)

    # state movement ----------------------------------------------------------
    def _stage(self, states: Sequence[HostState]):
        hP, hM, hV, hA = self.hP.numpy(), self.hM.numpy(), self.hV.numpy(), self.hA.numpy()
        hs, hk = self.hstep.numpy(), self.hkeys.numpy()
        for i, st in enumerate(states):
            o, n = self.p_off[i], self.n_params[i]
            hP[o:o + n] = st.params
            hM[o:o + n] = st.adam_m
            hV[o:o + n] = st.adam_v
            if self.a_sizes[i]:
                hA[self.a_off[i]:self.a_off[i] + self.a_sizes[i]] = np.asarray(st.A, np.float32).ravel()
            hs[4 * i] = st.step
            hk[8 * i:8 * i + 2] = st.model_key
            hk[8 * i + 4:8 * i + 6] = st.dataset_key

    def stage(self, states: Sequence[HostState]):
        """Packs the runs' host state into the pinned host arena (host-side only, no copy to the device)."""
        if len(states) != self.M:
            raise VaeB200Error("one HostState per run is required")
        self._stage(states)

    def upload(self, states: Optional[Sequence[HostState]] = None) -> int:
        """Host -> device copy of every model's state: ONE async copy of the pinned arena's state part on
        the current stream (``states`` given: staged into the pinned arena first).  Returns the bytes copied."""
        if states is not None:
            self.stage(states)
        n = self._state_words
        self.arena[:n].copy_(self.harena[:n], non_blocking=True)
        return n * 4

    def download(self, n_loss: int, states: Optional[Sequence[HostState]] = None) -> int:
        """Device -> host copy of the full state and losses[:n_loss] per model (one copy when n_loss is the
        whole chunk, else state + compacted losses).  Synchronises the current stream when ``states`` is
        given (values are unpacked into them)."""
        n = self._state_words
        if n_loss == self.max_chunk:
            self.harena.copy_(self.arena, non_blocking=True)
            nb = self.arena.numel() * 4
        else:
            self.harena[:n].copy_(self.arena[:n], non_blocking=True)
            nb = n * 4
            if n_loss > 0:
                src = self.loss.view(self.M, self.max_chunk)[:, :n_loss].contiguous()     # compact on the device
                self.hloss[:self.M * n_loss].copy_(src.reshape(-1), non_blocking=True)
                nb += self.M * n_loss * 4
        if states is not None:
            torch.cuda.current_stream().synchronize()
            self.unpack(states, n_loss)
        return nb

    def unpack(self, states: Sequence[HostState], n_loss: int):
        """Pinned host arena (after a completed download) -> the runs' HostState objects."""
        hP, hM, hV = self.hP.numpy(), self.hM.numpy(), self.hV.numpy()
        hs, hk = self.hstep.numpy(), self.hkeys.numpy()
        hl = self.hloss[:self.M * max(n_loss, 1)].view(self.M, max(n_loss, 1)).numpy()
        for i, st in enumerate(states):
            o, n = self.p_off[i], self.n_params[i]
            st.params = hP[o:o + n].copy()
            st.adam_m = hM[o:o + n].copy()
            st.adam_v = hV[o:o + n].copy()
            st.step = int(hs[4 * i])
            st.model_key = hk[8 * i:8 * i + 2].copy()
            st.dataset_key = hk[8 * i + 4:8 * i + 6].copy()
            if n_loss > 0:
                st.losses.append(hl[i, :n_loss].copy())

    # compute -----------------------------------------------------------------
    def prepare(self):
        """(Re)builds the launch plan: job validation, device job table upload and the launch-shaping
        environment knobs (VAE_SMALL_GM, VAE_SMALL_CLUSTER, VAE_MLP_CLUSTER) are all evaluated here, once."""
        self.plan = ops.pack_prepare(self.jobs, self.scratch)
        return self.plan

    def train_steps(self, n_steps: int):
        """n_steps iterations of the hot loop (model.py:221-222) for every packed model, in
        one persistent launch on the current stream (``vae_pack_train``: no per-job host work).
        Losses land in ``losses(i, n_steps)``."""
        if n_steps > self.max_chunk:
            raise VaeB200Error(f"n_steps={n_steps} exceeds max_chunk={self.max_chunk}")
        if self.plan is None:
            self.prepare()
        ops.pack_train(self.plan, n_steps)

    def train_host(self, states: Sequence[HostState], n_steps: int):
        """End-to-end: host state in, n_steps fused steps, host state + losses out."""
        h2d = self.upload(states)
        self.train_steps(n_steps)
        d2h = self.download(n_steps, states)
        return h2d, d2h


class OverlappedPack:
    """The end-to-end call with copies hidden behind compute: the M runs are cut into ``parts`` contiguous sub-packs
    (each a PackedRuns with its own device arena, pinned host arena, launch plan and CUDA stream).  ``train_host_all``
    enqueues, per part, [H2D state copy -> K fused steps -> D2H state + losses] on the part's stream: while one part's
    state is on the PCIe link the others' kernels run (a part of a GPU-filling pack still holds >= 1 CTA per two SMs,
    and the parts' kernels are co-resident), so a repeat costs about one copy + the kernel instead of copy + kernel +
    copy.  The current stream is joined on both sides: the call is ordered like a single-stream train_host.
    Results are bit-identical to one PackedRuns over the same runs (a model never sees its neighbours; eps sharing is
    decided per CTA on the device)."""

    def __init__(self, specs: Sequence[RunSpec], max_chunk: int, parts: int = 4, align: int = 4, device="cuda"):
        M = len(specs)
        parts = max(1, min(int(parts), (M + align - 1) // align))
        per = (M + parts - 1) // parts
        per = (per + align - 1) // align * align           # whole CTAs of `align` models stay together
        self.bounds = [(lo, min(M, lo + per)) for lo in range(0, M, per)]
        self.packs = [PackedRuns(specs[lo:hi], max_chunk, device) for lo, hi in self.bounds]
        self.streams = [torch.cuda.Stream(device=device) for _ in self.packs]
        self._fork = torch.cuda.Event()
        self._joins = [torch.cuda.Event() for _ in self.packs]
        self._graphs, self._warm = {}, {}
        self.M = M

    def stage(self, states: Sequence[HostState]):
        for (lo, hi), p in zip(self.bounds, self.packs):
            p.stage(states[lo:hi])

    def prepare(self):
        for p in self.packs:
            p.prepare()

    def _enqueue(self, n_steps: int):
        main = torch.cuda.current_stream()
        self._fork.record(main)
        h2d = d2h = 0
        for p, s, j in zip(self.packs, self.streams, self._joins):
            s.wait_event(self._fork)
            with torch.cuda.stream(s):
                h2d += p.upload()
                p.train_steps(n_steps)
                d2h += p.download(n_steps)
                j.record(s)
        for j in self._joins:
            main.wait_event(j)
        return h2d, d2h

    def train_host_all(self, n_steps: int, states: Optional[Sequence[HostState]] = None, graph: bool = True):
        """Staged host state in, n_steps fused steps, state + losses[:n_steps] back in the pinned host arenas
        (unpacked into ``states`` when given, which synchronises).  Returns (h2d bytes, d2h bytes).
        graph=True: the second call with the same n_steps captures the whole fork / copy / launch / copy / join pattern
        into a CUDA graph and later calls replay it (one host call per repeat instead of ~25: at 20-step launches
        the host, not the GPU, is otherwise what limits the rate)."""
        g = self._graphs.get(n_steps)
        if g is not None and graph:
            g[0].replay()
            h2d, d2h = g[1]
        elif graph and self._warm.get(n_steps):
            cg = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(cg):
                h2d, d2h = self._enqueue(n_steps)
            self._graphs[n_steps] = (cg, (h2d, d2h))
            cg.replay()
        else:
            h2d, d2h = self._enqueue(n_steps)
            self._warm[n_steps] = True
        if states is not None:
            torch.cuda.current_stream().synchronize()
            for (lo, hi), p in zip(self.bounds, self.packs):
                p.unpack(states[lo:hi], n_steps)
        return h2d, d2h

    def train_host_rounds(self, n_steps: int, rounds: int, graph: bool = True):
        """`rounds` consecutive end-to-end calls of every part, pipelined ACROSS rounds: each part's stream runs
        [H2D state -> n_steps fused steps -> D2H state + losses] `rounds` times back to back (a download leaves the
        advanced state in the pinned arena, so round r + 1 continues where round r stopped, exactly like repeated
        train_host calls), the parts are only joined at the very end.  In steady state the H2D engine, the D2H engine
        and the SMs all stay busy with different parts: the rate is set by the slowest of the three, not by their sum.
        The whole pattern is captured once per (n_steps, rounds) into a CUDA graph.  Returns (h2d, d2h) bytes per round."""
        key = (n_steps, rounds)
        g = self._graphs.get(key)
        if g is not None and graph:
            g[0].replay()
            return g[1]

        def enqueue():
            main = torch.cuda.current_stream()
            self._fork.record(main)
            h2d = d2h = 0
            for p, s, j in zip(self.packs, self.streams, self._joins):
                s.wait_event(self._fork)
                with torch.cuda.stream(s):
                    for _ in range(rounds):
                        a = p.upload()
                        p.train_steps(n_steps)
                        b = p.download(n_steps)
                    h2d += a
                    d2h += b
                    j.record(s)
            for j in self._joins:
                main.wait_event(j)
            return h2d, d2h
        if graph and self._warm.get(key):
            cg = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(cg):
                out = enqueue()
            self._graphs[key] = (cg, out)
            cg.replay()
            return out
        self._warm[key] = True
        return enqueue()

    def host_losses(self, n_steps: int) -> np.ndarray:
        """(M, n_steps) losses of the last train_host_all from the pinned host arenas (after a synchronize)."""
        out = []
        for p in self.packs:
            out.append(p.hloss[:p.M * n_steps].view(p.M, n_steps).numpy().copy())
        return np.concatenate(out, axis=0)
