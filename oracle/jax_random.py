"""ORACLE (test infrastructure) -- NumPy restatement of the slice of ``jax.random``
the reference's hot path uses.  See ``oracle/__init__.py`` for who may import this.

Third-party dependency restated: ``jax`` / ``jaxlib`` (un-vendored, un-pinned --
``README.md:8-13``; the API era is bounded by ``flax.nn`` / ``flax.optim`` use at
``networks.py:26,59`` and ``vae.py:113`` => the legacy, non-partitionable threefry
layout).  Reference call sites this file serves:

* ``random.PRNGKey``   -- ``model.py:29`` (model chain root, seed 0),
                          ``datasets.py:35`` (dataset chain root, ``-ds`` seed)
* ``random.split``     -- ``vae.py:56,125``, ``model.py:58,140,154,208``,
                          ``datasets.py:51``
* ``random.normal``    -- ``model.py:227`` (eps, shape (B, L+D)),
                          ``datasets.py:76,173,185,191,237,242``

Algorithm: Threefry-2x32 with 20 rounds (Random123), JAX's
"split the counter array in halves" lane layout, ``uniform`` via the mantissa
trick, ``normal = sqrt(2) * erf_inv(uniform(-1+ulp, 1))`` with Giles' single
precision ``erf_inv`` polynomial (XLA ``ErfInv32``), Horner steps fused.

Float determinism.  Every float operation below is a single IEEE-754 binary32 add / sub /
mul / sqrt or a FUSED multiply-add (``_fma``: one rounding), written out explicitly, so the
CUDA kernels (``__fadd_rn`` / ``__fmul_rn`` / ``__fmaf_rn`` / ``__fsqrt_rn``) and the C oracle
(explicit ``fmaf``, ``-ffp-contract=off``) reproduce these values bit for bit.  The Horner
steps of the polynomials are fused because XLA's CPU backend allows FP-op fusion
(``AllowFPOpFusion = Fast``) and the hosts have FMA units; whether the reference's executable
fused them cannot be observed here.  Both the fused and the unfused evaluation reproduce the
two documented values ``normal(PRNGKey(0))`` / ``normal(PRNGKey(42))`` bit for bit (they differ
from each other by <= 3 ulp on 4 % of draws), so the float path stays "<= a few ulp, parity
unpinned" either way.  The one place where XLA's own code generation cannot be restated
from the reference tree is ``log1p``; ``neg_log1p_neg`` below is our own <=1-ulp,
division-free binary32 evaluation (Cephes ``logf`` polynomial with an exact residual fold-in).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
F32 = np.float32

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All arguments uint32 (scalars or arrays); returns
    the two output lanes.  Random123 / ``jax._src.prng.threefry2x32`` [3P]."""
    with np.errstate(over="ignore"):
        k0 = U32(k0)
        k1 = U32(k1)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, U32(k0 ^ k1 ^ _PARITY))
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """``random.PRNGKey(seed)`` -> uint32[2] = [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def threefry_counts(key, counts) -> np.ndarray:
    """JAX's array form: pad to even length, first half -> lane 0 counters, second
    half -> lane 1 counters, output = concat(y0, y1)[:n]."""
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    return np.concatenate([y0, y1])[:n]


def random_bits(key, n: int) -> np.ndarray:
    """``_random_bits(key, 32, (n,))`` = threefry over iota(n) (n < 2**32)."""
    return threefry_counts(key, np.arange(n, dtype=U32))


def split(key, num: int = 2) -> np.ndarray:
    """``random.split(key, num)`` -> uint32[num, 2]."""
    return threefry_counts(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def fold_in(key, data: int) -> np.ndarray:
    return threefry_counts(key, np.array([0, data], dtype=U32))


# --------------------------------------------------------------------------- floats

_LO = np.nextafter(F32(-1.0), F32(0.0))        # -0.99999994
_SCALE = F32(F32(1.0) - _LO)                   # rounds to exactly 2.0f
assert _SCALE == F32(2.0)
_SQRT2 = F32(np.sqrt(2.0))

# Cephes logf polynomial, highest degree first.
_LOGP = [F32(c) for c in (7.0376836292e-2, -1.1514610310e-1, 1.1676998740e-1,
                          -1.2420140846e-1, 1.4249322787e-1, -1.6668057665e-1,
                          2.0000714765e-1, -2.4999993993e-1, 3.3333331174e-1)]
_LN2_HI = F32(0.693359375)
_LN2_LO = F32(-2.12194440e-4)

# Giles erf_inv, single precision (XLA ErfInv32), highest degree first.
_ERFINV_CENTRAL = [F32(c) for c in (2.81022636e-08, 3.43273939e-07, -3.5233877e-06,
                                    -4.39150654e-06, 0.00021858087, -0.00125372503,
                                    -0.00417768164, 0.246640727, 1.50140941)]
_ERFINV_TAIL = [F32(c) for c in (-0.000200214257, 0.000100950558, 0.00134934322,
                                 -0.00367342844, 0.00573950773, -0.0076224613,
                                 0.00943887047, 1.00167406, 2.83297682)]


def _fma(a, b, c) -> np.ndarray:
    """binary32 fused multiply-add.  The product of two binary32 values is exact in x87
    extended precision (64-bit significand); the sum is rounded there and then to binary32 --
    a double rounding that differs from a true FMA only when the extended sum lands exactly
    on a binary32 tie (probability ~2^-40 per operation).  The C oracle uses a true fmaf and is
    compared with this file bit for bit in tests/test_oracle_c.py."""
    ld = np.longdouble
    return (np.asarray(a, F32).astype(ld) * np.asarray(b, F32).astype(ld) + np.asarray(c, F32).astype(ld)).astype(F32)


def neg_log1p_neg(t) -> np.ndarray:
    """w = -log1p(-t) for binary32 t in [0, 1), using only binary32 add/mul.

    a = fl(1 - t); c = (1 - a) - t is the exact rounding residual (1 - t = a + c).
    a = m * 2**k with m in [sqrt(1/2), sqrt(2)); f = (m - 1) + c * 2**-k folds the
    residual back in where it matters (t small => k = 0, f = -t exactly);
    log(1 - t) = k*ln2 + log1p(f), log1p by the Cephes logf polynomial."""
    t = np.asarray(t, dtype=F32)
    one = F32(1.0)
    a = one - t
    c = (one - a) - t
    bits = a.view(U32)
    k = (bits >> U32(23)).astype(np.int32) - 127
    mbits = (bits & U32(0x007FFFFF)) | U32(0x3F800000)
    big = mbits >= U32(0x3FB504F3)                  # m >= sqrt(2)
    mbits = np.where(big, mbits - U32(0x00800000), mbits)
    k = np.where(big, k + 1, k)
    m = mbits.view(F32)
    inv_scale = ((127 - k).astype(np.int32) << 23).astype(U32).view(F32)   # 2**-k
    f = _fma(c, inv_scale, m - one)
    fk = k.astype(F32)
    z = f * f
    p = np.broadcast_to(_LOGP[0], f.shape)
    for coef in _LOGP[1:]:
        p = _fma(p, f, coef)
    y = (p * f) * z
    y = _fma(fk, _LN2_LO, y)
    y = _fma(F32(-0.5), z, y)
    r = f + y
    r = _fma(fk, _LN2_HI, r)
    return -r


def erf_inv(u) -> np.ndarray:
    """Single precision Giles erf_inv with unfused Horner steps."""
    u = np.asarray(u, dtype=F32)
    w = neg_log1p_neg(u * u)
    central = w < F32(5.0)
    wc = w - F32(2.5)
    with np.errstate(invalid="ignore"):
        wt = np.sqrt(np.where(central, F32(9.0), w)).astype(F32) - F32(3.0)
    ww = np.where(central, wc, wt)
    p = np.where(central, _ERFINV_CENTRAL[0], _ERFINV_TAIL[0])
    for cc, ct in zip(_ERFINV_CENTRAL[1:], _ERFINV_TAIL[1:]):
        p = _fma(p, ww, np.where(central, cc, ct))
    return (p * u).astype(F32)


def bits_to_unit(bits) -> np.ndarray:
    """uint32 bits -> float32 in [0, 1): bitcast((bits >> 9) | 0x3F800000) - 1."""
    bits = np.asarray(bits, dtype=U32)
    return ((bits >> U32(9)) | U32(0x3F800000)).view(F32) - F32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    f = bits_to_unit(random_bits(key, n))
    lo = F32(minval)
    scale = F32(F32(maxval) - lo)
    out = np.maximum(lo, f * scale + lo)
    return out.reshape(shape)


def normal_from_bits(bits) -> np.ndarray:
    f = bits_to_unit(bits)
    u = np.maximum(_LO, f * _SCALE + _LO)
    return (_SQRT2 * erf_inv(u)).astype(F32)


def normal(key, shape=()) -> np.ndarray:
    """``random.normal(key, shape)`` float32."""
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    return normal_from_bits(random_bits(key, n)).reshape(shape)


# ------------------------------------------------------------------ key-chain helpers

def model_chain_next(key):
    """Model-side convention (``vae.py:56,125``; ``model.py:140,154,208``):
    ``sub, self.key = split(self.key)`` -- row 0 is consumed, row 1 carried.
    Returns (consumed, carried)."""
    s = split(key)
    return s[0].copy(), s[1].copy()


def dataset_chain_next(key):
    """Dataset-side convention (``datasets.py:50-52``):
    ``self.key, key = split(self.key)`` -- row 0 carried, row 1 returned.
    Returns (returned, carried)."""
    s = split(key)
    return s[1].copy(), s[0].copy()
