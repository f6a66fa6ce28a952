"""Thin torch-plumbed wrappers over the C ABI (include/vae_b200.h).

torch is used for device memory and streams only; every function below enqueues
hand-written CUDA kernels from libvae_b200.so on torch's current stream and never
synchronises.  There is no CPU fallback: CPU tensors are rejected.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import VaeB200Error, VaeDataDesc, VaeDesc, VaePackPlan, VaeTrainJob, check


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor], dtype=None, what="tensor") -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise VaeB200Error(f"{what} must be a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise VaeB200Error(f"{what} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise VaeB200Error(f"{what} must have dtype {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


def key_tensor(key, device="cuda") -> torch.Tensor:
    """uint32[2] PRNG key on the device (jax.random legacy key layout)."""
    arr = np.asarray(key, dtype=np.uint32).reshape(2)
    return torch.from_numpy(arr.copy()).to(device)


def prng_key(seed: int) -> np.ndarray:
    """random.PRNGKey(seed) (model.py:29, datasets.py:35): [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


# ---------------------------------------------------------------------- jax.random

def random_split(key: torch.Tensor, num: int = 2) -> torch.Tensor:
    """random.split(key, num) -> uint32[num, 2]."""
    out = torch.empty((num, 2), dtype=torch.uint32, device=key.device)
    check(_lib.load().vae_random_split(_ptr(key, torch.uint32, "key"), _ptr(out), num, _stream()), "vae_random_split")
    return out


def random_bits(key: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.empty((n,), dtype=torch.uint32, device=key.device)
    check(_lib.load().vae_random_bits(_ptr(key, torch.uint32, "key"), _ptr(out), n, _stream()), "vae_random_bits")
    return out


def random_normal(key: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """random.normal(key, shape) float32."""
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    out = torch.empty((n,), dtype=torch.float32, device=key.device)
    check(_lib.load().vae_random_normal(_ptr(key, torch.uint32, "key"), _ptr(out), n, _stream()), "vae_random_normal")
    return out.reshape(shape)


def random_normal_rows(key: torch.Tensor, rows: int, cols: int, row_begin: int, row_end: int) -> torch.Tensor:
    """Rows [row_begin, row_end) of random.normal(key, (rows, cols))."""
    out = torch.empty((row_end - row_begin, cols), dtype=torch.float32, device=key.device)
    check(_lib.load().vae_random_normal_rows(_ptr(key, torch.uint32, "key"), _ptr(out), rows, cols,
                                              row_begin, row_end, _stream()), "vae_random_normal_rows")
    return out


def random_normal_pairs(key: torch.Tensor, rows: int, cols: int, lo: int, hi: int) -> torch.Tensor:
    """Rows [lo, hi) and partners [rows/2 + lo, rows/2 + hi) of random.normal(key, (rows, cols))."""
    out = torch.empty((2 * (hi - lo), cols), dtype=torch.float32, device=key.device)
    check(_lib.load().vae_random_normal_pairs(_ptr(key, torch.uint32, "key"), _ptr(out), rows, cols, lo, hi,
                                               _stream()), "vae_random_normal_pairs")
    return out


# ------------------------------------------------------------------------ datasets

def dataset_batch_pairs(data: VaeDataDesc, key: torch.Tensor, A: Optional[torch.Tensor], size: int, lo: int,
                        hi: int) -> torch.Tensor:
    lib = _lib.load()
    width = lib.vae_data_width(C.byref(data))
    out = torch.empty((2 * (hi - lo), width), dtype=torch.float32, device=key.device)
    check(lib.vae_dataset_batch_pairs(C.byref(data), _ptr(key, torch.uint32, "key"), _ptr(A, torch.float32, "A"),
                                      size, lo, hi, _ptr(out), _stream()), "vae_dataset_batch_pairs")
    return out


def dataset_batch(data: VaeDataDesc, key: torch.Tensor, A: Optional[torch.Tensor], size: int) -> torch.Tensor:
    lib = _lib.load()
    width = lib.vae_data_width(C.byref(data))
    out = torch.empty((size, width), dtype=torch.float32, device=key.device)
    check(lib.vae_dataset_batch(C.byref(data), _ptr(key, torch.uint32, "key"), _ptr(A, torch.float32, "A"),
                                size, _ptr(out), _stream()), "vae_dataset_batch")
    return out


def dataset_score(data: VaeDataDesc, A: Optional[torch.Tensor], batch: torch.Tensor) -> torch.Tensor:
    out = torch.empty((2,), dtype=torch.float32, device=batch.device)
    check(_lib.load().vae_dataset_score(C.byref(data), _ptr(A, torch.float32, "A"),
                                        _ptr(batch, torch.float32, "batch"), batch.shape[0], _ptr(out),
                                        _stream()), "vae_dataset_score")
    return out


# ------------------------------------------------------------------------ VAE step

def param_count(desc: VaeDesc) -> int:
    n = _lib.load().vae_param_count(C.byref(desc))
    if n < 0:
        check(int(n), "vae_param_count")
    return int(n)


class Workspace:
    """Caller-owned device scratch for the per-op step path, grown on demand."""

    def __init__(self, device="cuda"):
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, desc: VaeDesc, rows: int):
        need = int(_lib.load().vae_workspace_bytes(C.byref(desc), rows))
        if need == 0:
            raise VaeB200Error("vae_workspace_bytes returned 0: invalid model description")
        if self.buf is None or self.buf.numel() < need:
            self.buf = torch.empty((need,), dtype=torch.uint8, device=self.device)
        return C.c_void_p(self.buf.data_ptr()), C.c_size_t(self.buf.numel())


def train_step(desc: VaeDesc, params, adam_m, adam_v, step, x, z, loss, ws: Workspace):
    """VAE.train_step (networks.py:87-101): params / adam_m / adam_v / step updated in place,
    loss (float32[1]) written."""
    p, n = ws.get(desc, desc.batch)
    check(_lib.load().vae_train_step(C.byref(desc), _ptr(params, torch.float32, "params"),
                                     _ptr(adam_m, torch.float32, "adam_m"), _ptr(adam_v, torch.float32, "adam_v"),
                                     _ptr(step, torch.int32, "step"), _ptr(x, torch.float32, "x"),
                                     _ptr(z, torch.float32, "z"), _ptr(loss, torch.float32, "loss"), p, n,
                                     _stream()), "vae_train_step")


def loss_and_grad(desc: VaeDesc, params, x, z, rows: int, grad_scale_rows: int, loss, grad, ws: Workspace):
    p, n = ws.get(desc, rows)
    check(_lib.load().vae_loss_and_grad(C.byref(desc), _ptr(params, torch.float32, "params"),
                                        _ptr(x, torch.float32, "x"), _ptr(z, torch.float32, "z"), rows,
                                        grad_scale_rows, _ptr(loss, torch.float32, "loss"),
                                        _ptr(grad, torch.float32, "grad"), p, n, _stream()), "vae_loss_and_grad")


def adam_update(desc: VaeDesc, params, adam_m, adam_v, step, grad):
    check(_lib.load().vae_adam_update(C.byref(desc), _ptr(params, torch.float32, "params"),
                                      _ptr(adam_m, torch.float32, "adam_m"), _ptr(adam_v, torch.float32, "adam_v"),
                                      _ptr(step, torch.int32, "step"), _ptr(grad, torch.float32, "grad"),
                                      _stream()), "vae_adam_update")


def grad_segments(desc: VaeDesc):
    """[(begin, end)] ranges of the flat gradient in the order the backward pass completes them."""
    lib = _lib.load()
    n = int(lib.vae_grad_segments(C.byref(desc), None, 0))
    if n < 0:
        check(n, "vae_grad_segments")
    buf = (C.c_int64 * (2 * n))()
    check(min(0, int(lib.vae_grad_segments(C.byref(desc), buf, n))), "vae_grad_segments")
    return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)]


def loss_and_grad_events(desc: VaeDesc, params, x, z, rows: int, grad_scale_rows: int, loss, grad, ws: Workspace, events):
    """vae_loss_and_grad_events: `events` = one recorded-at-least-once torch.cuda.Event per gradient segment."""
    p, n = ws.get(desc, rows)
    arr = (C.c_void_p * len(events))(*[C.c_void_p(e.cuda_event) for e in events])
    check(_lib.load().vae_loss_and_grad_events(C.byref(desc), _ptr(params, torch.float32, "params"),
                                               _ptr(x, torch.float32, "x"), _ptr(z, torch.float32, "z"), rows,
                                               grad_scale_rows, _ptr(loss, torch.float32, "loss"),
                                               _ptr(grad, torch.float32, "grad"), p, n, arr, _stream()),
          "vae_loss_and_grad_events")


def adam_update_range(desc: VaeDesc, params, adam_m, adam_v, step, grad, begin: int, end: int):
    check(_lib.load().vae_adam_update_range(C.byref(desc), _ptr(params, torch.float32, "params"),
                                            _ptr(adam_m, torch.float32, "adam_m"), _ptr(adam_v, torch.float32, "adam_v"),
                                            _ptr(step, torch.int32, "step"), _ptr(grad, torch.float32, "grad"),
                                            int(begin), int(end), _stream()), "vae_adam_update_range")


def adam_step_increment(step):
    check(_lib.load().vae_adam_step_increment(_ptr(step, torch.int32, "step"), _stream()), "vae_adam_step_increment")


def vae_loss(desc: VaeDesc, params, x, z, rows: int, ws: Workspace) -> torch.Tensor:
    """VAE.loss (networks.py:103-113): float32[3] = loss, dkl.mean, mse.mean."""
    out = torch.empty((3,), dtype=torch.float32, device=params.device)
    p, n = ws.get(desc, rows)
    check(_lib.load().vae_loss(C.byref(desc), _ptr(params, torch.float32, "params"), _ptr(x, torch.float32, "x"),
                               _ptr(z, torch.float32, "z"), rows, _ptr(out), p, n, _stream()), "vae_loss")
    return out


def vae_sample(desc: VaeDesc, params, z, rows: int, current_epsilon, ws: Workspace) -> torch.Tensor:
    """Sampling-mode forward (networks.py:62-65 via vae.py:191-201)."""
    out = torch.empty((rows, desc.data_dim), dtype=torch.float32, device=params.device)
    p, n = ws.get(desc, rows)
    check(_lib.load().vae_sample(C.byref(desc), _ptr(params, torch.float32, "params"), _ptr(z, torch.float32, "z"),
                                 rows, _ptr(current_epsilon, torch.float32, "current_epsilon"), _ptr(out), p, n,
                                 _stream()), "vae_sample")
    return out


# ------------------------------------------------------------------- fused training

def make_job(desc: VaeDesc, data: VaeDataDesc, params, adam_m, adam_v, step, model_key, dataset_key, A,
             loss_out, workspace=None) -> VaeTrainJob:
    """The job records raw device pointers; the caller keeps every tensor alive (and
    unmoved) until the launches that use the job have completed."""
    j = VaeTrainJob()
    j.desc = desc
    j.data = data
    j.params = _ptr(params, torch.float32, "params").value
    j.adam_m = _ptr(adam_m, torch.float32, "adam_m").value
    j.adam_v = _ptr(adam_v, torch.float32, "adam_v").value
    j.step = _ptr(step, torch.int32, "step").value
    j.model_key = _ptr(model_key, torch.uint32, "model_key").value
    j.dataset_key = _ptr(dataset_key, torch.uint32, "dataset_key").value
    j.A = _ptr(A, torch.float32, "A").value
    j.loss_out = _ptr(loss_out, torch.float32, "loss_out").value
    j.workspace = _ptr(workspace).value
    return j


def fused_supported(job: VaeTrainJob) -> int:
    """0: no fused kernel; 1: single-CTA kernel; 2: thread-block-cluster kernel (needs a workspace)."""
    return int(_lib.load().vae_train_fused_supported(C.byref(job)))


def fused_workspace(job: VaeTrainJob, device="cuda") -> Optional[torch.Tensor]:
    """Device workspace a class-2 job needs (None when the kernel needs none)."""
    n = int(_lib.load().vae_train_fused_workspace_bytes(C.byref(job)))
    return torch.zeros((n,), dtype=torch.uint8, device=device) if n > 0 else None


def fused_ctas_per_sm(job: VaeTrainJob) -> int:
    n = int(_lib.load().vae_train_fused_ctas_per_sm(C.byref(job)))
    if n < 0:
        check(n, "vae_train_fused_ctas_per_sm")
    return n


def fused_resident_jobs(job: VaeTrainJob) -> int:
    n = int(_lib.load().vae_train_fused_resident_jobs(C.byref(job)))
    if n < 0:
        check(n, "vae_train_fused_resident_jobs")
    return n


def train_fused(jobs: Sequence[VaeTrainJob], n_steps: int, scratch: torch.Tensor):
    """One persistent launch: n_steps consecutive hot-loop iterations (model.py:221-222)
    for every job."""
    n = len(jobs)
    arr = (VaeTrainJob * n)(*jobs)
    check(_lib.load().vae_train_fused(arr, n, int(n_steps), _ptr(scratch), C.c_size_t(scratch.numel() * scratch.element_size()),
                                      _stream()), "vae_train_fused")


def pack_prepare(jobs: Sequence[VaeTrainJob], scratch: torch.Tensor) -> VaePackPlan:
    """Validates the jobs and uploads their device job table ONCE (vae_pack_prepare); the returned
    plan makes every later ``pack_train`` a single kernel launch with no per-job host work.  The
    caller keeps ``scratch`` and every tensor the jobs point to alive while the plan is in use."""
    n = len(jobs)
    arr = (VaeTrainJob * n)(*jobs)
    plan = VaePackPlan()
    check(_lib.load().vae_pack_prepare(arr, n, _ptr(scratch), C.c_size_t(scratch.numel() * scratch.element_size()),
                                       C.byref(plan), _stream()), "vae_pack_prepare")
    return plan


def pack_train(plan: VaePackPlan, n_steps: int):
    """n_steps hot-loop iterations (model.py:221-222) of every job of the prepared pack: ONE launch."""
    check(_lib.load().vae_pack_train(C.byref(plan), int(n_steps), _stream()), "vae_pack_train")


def fused_scratch(n_jobs: int, device="cuda") -> torch.Tensor:
    need = int(_lib.load().vae_train_fused_scratch_bytes(n_jobs))
    return torch.empty((max(need, 16),), dtype=torch.uint8, device=device)


# ------------------------------------------------------------ tensor-core building blocks

def split_tf32(x: torch.Tensor):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(_lib.load().vae_split_tf32(_ptr(x, torch.float32, "x"), _ptr(hi), _ptr(lo), x.numel(), _stream()), "vae_split_tf32")
    return hi, lo


def transpose_split_tf32(x: torch.Tensor):
    rows, cols = x.shape
    hi = torch.empty((cols, rows), dtype=torch.float32, device=x.device)
    lo = torch.empty_like(hi)
    check(_lib.load().vae_transpose_split_tf32(_ptr(x, torch.float32, "x"), cols, _ptr(hi), _ptr(lo), rows, rows, cols,
                                               _stream()), "vae_transpose_split_tf32")
    return hi, lo


def gemm_nt_3xtf32(Ah, Al, Bh, Bl, bias=None, relu=False, mask=None, out=None) -> torch.Tensor:
    """C[M,N] = epi((Ah+Al)[M,K] @ (Bh+Bl)[N,K]^T) on the tensor cores (3xTF32)."""
    M, K = Ah.shape
    N = Bh.shape[0]
    C_ = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=Ah.device)
    flags = (1 if bias is not None else 0) | (2 if relu else 0) | (4 if mask is not None else 0)
    check(_lib.load().vae_gemm_nt_3xtf32(M, N, K, _ptr(Ah, torch.float32, "Ah"), _ptr(Al, torch.float32, "Al"), K,
                                         _ptr(Bh, torch.float32, "Bh"), _ptr(Bl, torch.float32, "Bl"), K, _ptr(C_), N,
                                         _ptr(bias, torch.float32, "bias"), _ptr(mask, torch.float32, "mask"), N, flags,
                                         _stream()), "vae_gemm_nt_3xtf32")
    return C_


def gemm_tn_3xtf32(Ath, Atl, Bth, Btl, out=None) -> torch.Tensor:
    """C[M,N] = (Ath+Atl)[K,M]^T @ (Bth+Btl)[K,N] on the tensor cores (3xTF32, MN-major operands)."""
    K, M = Ath.shape
    N = Bth.shape[1]
    C_ = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=Ath.device)
    check(_lib.load().vae_gemm_tn_3xtf32(M, N, K, _ptr(Ath, torch.float32, "Ath"), _ptr(Atl, torch.float32, "Atl"), M,
                                         _ptr(Bth, torch.float32, "Bth"), _ptr(Btl, torch.float32, "Btl"), N, _ptr(C_), N,
                                         _stream()), "vae_gemm_tn_3xtf32")
    return C_


def ffma_probe(blocks: int, threads: int, iters: int, sink: torch.Tensor) -> float:
    flops = C.c_double(0.0)
    check(_lib.load().vae_ffma_probe(_ptr(sink, torch.float32, "sink"), blocks, threads, iters, C.byref(flops),
                                     _stream()), "vae_ffma_probe")
    return flops.value
