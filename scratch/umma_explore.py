import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_umma_probe import sw128_image, trunc_tf32, run, relerr
from vae_training_b200 import _lib as lib

def img_base32(mat):
    rows, cols = mat.shape
    img = np.zeros((cols // 32, rows, 32), np.float32)
    for a in range(cols // 32):
        blk = mat[:, 32*a:32*a+32].reshape(rows, 4, 8)
        for r in range(rows):
            perm = np.arange(4) ^ (r % 4)
            img[a, r].reshape(4, 8)[perm] = blk[r]
    return img.reshape(-1)

def img_none(mat):
    return None

def desc(lbo, sbo, lt):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (lt << 61)

def idesc(n, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | (8 << 24)

rng = np.random.default_rng(2)
k, n = 32, 104
A = trunc_tf32(rng.standard_normal((128, k)).astype(np.float32))
W = trunc_tf32(rng.standard_normal((k, 128)).astype(np.float32))
ref = A.astype(np.float64) @ W.astype(np.float64)[:, :n]
KA = desc(16, 1024, 2)
for name, img, lt, lbo, sbo, bstep in [
    ("B MN sw128 lbo=atom sbo=1024", sw128_image(W), 2, k*128, 1024, 1024),
    ("B MN sw128 lbo=1024 sbo=atom", sw128_image(W), 2, 1024, k*128, 1024),
    ("B MN base32 lbo=atom sbo=512", img_base32(W), 1, k*128, 512, 1024),
    ("B MN base32 lbo=atom sbo=1024", img_base32(W), 1, k*128, 1024, 1024),
    ("B MN base32 lbo=512 sbo=atom", img_base32(W), 1, 512, k*128, 1024),
]:
    try:
        got = run(lib, sw128_image(A), img, KA, desc(lbo, sbo, lt), idesc(n, False, True), 4, 32, bstep, n)
        print(name, "relerr", relerr(got, ref), "nonzero", np.count_nonzero(got), "nan", np.isnan(got).sum())
    except Exception as e:
        print(name, "EXC", e)
        break
