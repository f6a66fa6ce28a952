"""ctypes binding of libvae_b200.so (the C ABI declared in include/vae_b200.h).

There is no CPU fallback: if the shared library is missing the import of any
product module raises, and every entry point raises ``VaeB200Error`` on a
non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAE_B200_LIB") or os.path.join(HERE, "libvae_b200.so")      # override: A/B builds
VAE_MAX_HIDDEN = 8

DATA_KINDS = {"sphere": 0, "linear_gaussian": 1, "sigmoid": 2}


class VaeB200Error(RuntimeError):
    pass


class VaeDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("data_dim", C.c_int32), ("latent_dim", C.c_int32),
        ("n_enc_hidden", C.c_int32), ("enc_hidden", C.c_int32 * VAE_MAX_HIDDEN),
        ("n_dec_hidden", C.c_int32), ("dec_hidden", C.c_int32 * VAE_MAX_HIDDEN),
        ("sig_decoder", C.c_int32), ("tunable_decoder_var", C.c_int32), ("reserved_", C.c_int32),
        ("epsilon0", C.c_double), ("lr", C.c_double), ("beta1", C.c_double),
        ("beta2", C.c_double), ("adam_eps", C.c_double),
    ]


class VaeDataDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("padding_dim", C.c_int32),
                ("intrinsic_dim", C.c_int32)]


class VaeTrainJob(C.Structure):
    _fields_ = [
        ("desc", VaeDesc), ("data", VaeDataDesc),
        ("params", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("step", C.c_void_p), ("model_key", C.c_void_p), ("dataset_key", C.c_void_p),
        ("A", C.c_void_p), ("loss_out", C.c_void_p), ("workspace", C.c_void_p),
    ]


class VaePackPlan(C.Structure):
    """Caller-owned launch plan of a prepared pack (include/vae_b200.h: VaePackPlan)."""
    _fields_ = [
        ("abi", C.c_int32), ("cls", C.c_int32), ("n_jobs", C.c_int32), ("shape_id", C.c_int32),
        ("homogeneous", C.c_int32), ("gm", C.c_int32), ("cluster", C.c_int32), ("nbuf", C.c_int32),
        ("smem_bytes", C.c_uint64), ("d_jobs", C.c_void_p), ("device", C.c_int32), ("tc", C.c_int32), ("reserved_", C.c_int32 * 4),
    ]


_P = C.c_void_p
_PROTOTYPES = {
    "vae_param_count": (C.c_int64, [C.POINTER(VaeDesc)]),
    "vae_data_width": (C.c_int32, [C.POINTER(VaeDataDesc)]),
    "vae_data_normals_per_row": (C.c_int32, [C.POINTER(VaeDataDesc)]),
    "vae_random_split": (C.c_int, [_P, _P, C.c_int32, _P]),
    "vae_random_bits": (C.c_int, [_P, _P, C.c_int64, _P]),
    "vae_random_normal": (C.c_int, [_P, _P, C.c_int64, _P]),
    "vae_random_normal_rows": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    "vae_random_normal_pairs": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    "vae_dataset_batch_pairs": (C.c_int, [C.POINTER(VaeDataDesc), _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "vae_dataset_batch": (C.c_int, [C.POINTER(VaeDataDesc), _P, _P, C.c_int32, _P, _P]),
    "vae_dataset_score": (C.c_int, [C.POINTER(VaeDataDesc), _P, _P, C.c_int32, _P, _P]),
    "vae_workspace_bytes": (C.c_size_t, [C.POINTER(VaeDesc), C.c_int32]),
    "vae_train_step": (C.c_int, [C.POINTER(VaeDesc), _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vae_loss_and_grad": (C.c_int, [C.POINTER(VaeDesc), _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P,
                                    C.c_size_t, _P]),
    "vae_adam_update": (C.c_int, [C.POINTER(VaeDesc), _P, _P, _P, _P, _P, _P]),
    "vae_grad_segments": (C.c_int32, [C.POINTER(VaeDesc), _P, C.c_int32]),
    "vae_loss_and_grad_events": (C.c_int, [C.POINTER(VaeDesc), _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P,
                                           C.c_size_t, _P, _P]),
    "vae_adam_update_range": (C.c_int, [C.POINTER(VaeDesc), _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P]),
    "vae_adam_step_increment": (C.c_int, [_P, _P]),
    "vae_loss": (C.c_int, [C.POINTER(VaeDesc), _P, _P, _P, C.c_int32, _P, _P, C.c_size_t, _P]),
    "vae_sample": (C.c_int, [C.POINTER(VaeDesc), _P, _P, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "vae_train_fused_supported": (C.c_int, [C.POINTER(VaeTrainJob)]),
    "vae_train_fused_workspace_bytes": (C.c_size_t, [C.POINTER(VaeTrainJob)]),
    "vae_train_fused_ctas_per_sm": (C.c_int, [C.POINTER(VaeTrainJob)]),
    "vae_train_fused_resident_jobs": (C.c_int, [C.POINTER(VaeTrainJob)]),
    "vae_train_fused_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "vae_train_fused": (C.c_int, [C.POINTER(VaeTrainJob), C.c_int32, C.c_int64, _P, C.c_size_t, _P]),
    "vae_pack_prepare": (C.c_int, [C.POINTER(VaeTrainJob), C.c_int32, _P, C.c_size_t, C.POINTER(VaePackPlan), _P]),
    "vae_pack_train": (C.c_int, [C.POINTER(VaePackPlan), C.c_int64, _P]),
    "vae_split_tf32": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "vae_transpose_split_tf32": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "vae_gemm_tn_3xtf32": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int64, _P, _P, C.c_int64, _P, C.c_int64, _P]),
    "vae_gemm_nt_3xtf32": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int64, _P, _P, C.c_int64, _P,
                                     C.c_int64, _P, _P, C.c_int64, C.c_int32, _P]),
    "vae_umma_probe": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "vae_last_error": (C.c_char_p, []),
    "vae_abi_version": (C.c_int, []),
    "vae_launch_count": (C.c_int64, []),
    "vae_ffma_probe": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), _P]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def load():
    """Loads libvae_b200.so; raises VaeB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VaeB200Error(
            f"{LIB_PATH} is missing: build it with `python -m vae_training_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vae_last_error().decode("utf-8", "replace")
        raise VaeB200Error(f"{what or 'libvae_b200'} failed (rc={rc}): {msg}")


def make_desc(batch, data_dim, latent_dim, enc_hidden=(), dec_hidden=(), sig_decoder=False,
              tunable_decoder_var=True, epsilon0=0.0, lr=1e-4, beta1=0.9, beta2=0.999,
              adam_eps=1e-8) -> VaeDesc:
    if len(enc_hidden) > VAE_MAX_HIDDEN or len(dec_hidden) > VAE_MAX_HIDDEN:
        raise VaeB200Error(f"at most {VAE_MAX_HIDDEN} hidden layers are supported")
    d = VaeDesc()
    d.batch, d.data_dim, d.latent_dim = int(batch), int(data_dim), int(latent_dim)
    d.n_enc_hidden = len(enc_hidden)
    d.n_dec_hidden = len(dec_hidden)
    for i, h in enumerate(enc_hidden):
        d.enc_hidden[i] = int(h)
    for i, h in enumerate(dec_hidden):
        d.dec_hidden[i] = int(h)
    d.sig_decoder = int(bool(sig_decoder))
    d.tunable_decoder_var = int(bool(tunable_decoder_var))
    d.epsilon0, d.lr, d.beta1, d.beta2, d.adam_eps = epsilon0, lr, beta1, beta2, adam_eps
    return d


def make_data_desc(kind, dim, padding_dim=0, intrinsic_dim=3) -> VaeDataDesc:
    k = DATA_KINDS[kind] if isinstance(kind, str) else int(kind)
    return VaeDataDesc(k, int(dim), int(padding_dim), int(intrinsic_dim))
