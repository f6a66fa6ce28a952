"""ORACLE (test infrastructure) -- ctypes loader for the plain-C restatement in
``oracle/c/`` (the timed CPU baseline).  See ``oracle/__init__.py`` for who may import
this: tests/, ``__graft_entry__.smoke()/build()`` and bench.py's CPU legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(HERE, "c")
LIB = os.path.join(CDIR, "libvae_oracle.so")
MAXH = 8
KINDS = {"sphere": 0, "linear_gaussian": 1, "sigmoid": 2}


class OrDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("data_dim", C.c_int32), ("latent_dim", C.c_int32),
                ("n_enc_hidden", C.c_int32), ("enc_hidden", C.c_int32 * MAXH),
                ("n_dec_hidden", C.c_int32), ("dec_hidden", C.c_int32 * MAXH),
                ("sig_decoder", C.c_int32), ("tunable_decoder_var", C.c_int32), ("reserved_", C.c_int32),
                ("epsilon0", C.c_double), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("adam_eps", C.c_double)]


class OrData(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("padding_dim", C.c_int32), ("intrinsic_dim", C.c_int32)]


class OrJob(C.Structure):
    _fields_ = [("desc", OrDesc), ("data", OrData), ("params", C.c_void_p), ("adam_m", C.c_void_p),
                ("adam_v", C.c_void_p), ("step", C.c_void_p), ("model_key", C.c_void_p),
                ("dataset_key", C.c_void_p), ("A", C.c_void_p), ("loss_out", C.c_void_p)]


def _cpu_has_avx2_fma() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
        return " avx2" in flags and " fma" in flags
    except OSError:
        return False


def build(force: bool = False) -> str:
    srcs = [os.path.join(CDIR, f) for f in ("or_rng.c", "or_step.c", "vae_oracle.h", "Makefile")]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    if force or stale:
        simd = "1" if _cpu_has_avx2_fma() else "0"
        proc = subprocess.run(["make", "-C", CDIR, "-B", f"SIMD={simd}", "CC=gcc"], capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("building oracle/c failed:\n" + proc.stdout + proc.stderr)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.or_param_count.restype = C.c_int64
        lib.or_loss_and_grad.restype = C.c_float
        lib.or_data_width.restype = C.c_int32
        _lib = lib
    return _lib


def make_desc(d) -> OrDesc:
    """From an oracle.vae_step.VaeDesc."""
    o = OrDesc()
    o.batch, o.data_dim, o.latent_dim = d.batch, d.data_dim, d.latent_dim
    o.n_enc_hidden, o.n_dec_hidden = len(d.enc_hidden), len(d.dec_hidden)
    for i, h in enumerate(d.enc_hidden):
        o.enc_hidden[i] = h
    for i, h in enumerate(d.dec_hidden):
        o.dec_hidden[i] = h
    o.sig_decoder, o.tunable_decoder_var = int(d.sig_decoder), int(d.tunable_decoder_var)
    o.epsilon0, o.lr, o.beta1, o.beta2, o.adam_eps = d.epsilon0, d.lr, d.beta1, d.beta2, d.adam_eps
    return o


def make_data(name, dim, padding_dim, intrinsic_dim=3) -> OrData:
    return OrData(KINDS[name], dim, padding_dim, intrinsic_dim)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def random_bits(key, n):
    out = np.empty(n, np.uint32)
    load().or_random_bits(_p(np.ascontiguousarray(key, np.uint32)), _p(out), C.c_int64(n))
    return out


def split(key, num=2):
    out = np.empty((num, 2), np.uint32)
    load().or_split(_p(np.ascontiguousarray(key, np.uint32)), _p(out), num)
    return out


def normal(key, shape):
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    out = np.empty(n, np.float32)
    load().or_normal(_p(np.ascontiguousarray(key, np.uint32)), _p(out), C.c_int64(n))
    return out.reshape(shape)


def dataset_batch(data: OrData, key, A, size):
    lib = load()
    out = np.empty((size, lib.or_data_width(C.byref(data))), np.float32)
    A = None if A is None else np.ascontiguousarray(A, np.float32)
    lib.or_dataset_batch(C.byref(data), _p(np.ascontiguousarray(key, np.uint32)), _p(A), size, _p(out))
    return out


def loss_and_grad(desc: OrDesc, params, x, z):
    lib = load()
    grad = np.zeros(lib.or_param_count(C.byref(desc)), np.float32)
    loss = lib.or_loss_and_grad(C.byref(desc), _p(np.ascontiguousarray(params, np.float32)),
                                _p(np.ascontiguousarray(x, np.float32)), _p(np.ascontiguousarray(z, np.float32)),
                                _p(grad))
    return float(loss), grad


class Job:
    """Host-side state of one model for or_train_loop(_mt); arrays are owned here."""

    def __init__(self, desc: OrDesc, data: OrData, params, model_key, dataset_key, A, n_steps):
        self.params = np.ascontiguousarray(params, np.float32).copy()
        self.m = np.zeros_like(self.params)
        self.v = np.zeros_like(self.params)
        self.step = np.zeros(1, np.int32)
        self.model_key = np.ascontiguousarray(model_key, np.uint32).copy()
        self.dataset_key = np.ascontiguousarray(dataset_key, np.uint32).copy()
        self.A = None if A is None else np.ascontiguousarray(A, np.float32).copy()
        self.loss = np.zeros(max(int(n_steps), 1), np.float32)
        self.c = OrJob(desc, data, _p(self.params).value, _p(self.m).value, _p(self.v).value, _p(self.step).value,
                       _p(self.model_key).value, _p(self.dataset_key).value, _p(self.A).value, _p(self.loss).value)


def train_loop(jobs, n_steps, n_threads=1):
    arr = (OrJob * len(jobs))(*[j.c for j in jobs])
    rc = load().or_train_loop_mt(arr, len(jobs), C.c_int64(n_steps), n_threads)
    if rc != 0:
        raise RuntimeError("or_train_loop_mt failed")
