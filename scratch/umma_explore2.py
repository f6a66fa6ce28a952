import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_umma_probe import trunc_tf32, run, relerr
from vae_training_b200 import _lib as lib

def img_core(mat):
    """[rows][cols] -> [g = rows/8][c = cols/4][8][4] (no swizzle; core matrix = 8 rows x 16 B, contiguous 128 B)"""
    rows, cols = mat.shape
    return np.ascontiguousarray(mat.reshape(rows // 8, 8, cols // 4, 4).transpose(0, 2, 1, 3)).reshape(-1).astype(np.float32)

def desc(lbo, sbo, lt=0):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (lt << 61)

def idesc(n, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | (8 << 24)

rng = np.random.default_rng(3)
n = 104
# 1. K-major A [128][K=40], K-major B [N=104][K=40]
K = 40
A = trunc_tf32(rng.standard_normal((128, K)).astype(np.float32)); B = trunc_tf32(rng.standard_normal((n, K)).astype(np.float32))
GSa = K * 32
got = run(lib, img_core(A), img_core(B), desc(128, GSa), desc(128, GSa), idesc(n, 0, 0), K // 8, 256, 256, n)
print("K-major A, K-major B (interleave):", relerr(got, A.astype(np.float64) @ B.astype(np.float64).T))
# 2. K-major A [128][K=40], MN-major B from W[k=40][n=104]
W = trunc_tf32(rng.standard_normal((K, n)).astype(np.float32))
GSw = n * 32
got = run(lib, img_core(A), img_core(W), desc(128, GSa), desc(GSw, 128), idesc(n, 0, 1), K // 8, 256, GSw, n)
print("K-major A, MN-major B (interleave):", relerr(got, A.astype(np.float64) @ W.astype(np.float64)))
# 3. MN-major A from X[b=24][f=128], MN-major B from dY[b=24][n=104]
X = trunc_tf32(rng.standard_normal((24, 128)).astype(np.float32)); dY = trunc_tf32(rng.standard_normal((24, n)).astype(np.float32))
got = run(lib, img_core(X), img_core(dY), desc(128 * 32, 128), desc(n * 32, 128), idesc(n, 1, 1), 3, 128 * 32, n * 32, n)
print("MN-major A, MN-major B (interleave):", relerr(got, X.astype(np.float64).T @ dY.astype(np.float64)))
# 4. same as 3 with LBO/SBO swapped (in case the convention is the other way round)
got = run(lib, img_core(X), img_core(dY), desc(128, 128 * 32), desc(128, n * 32), idesc(n, 1, 1), 3, 128 * 32, n * 32, n)
print("MN-major A, MN-major B (interleave, swapped):", relerr(got, X.astype(np.float64).T @ dY.astype(np.float64)))
got = run(lib, img_core(A), img_core(W), desc(128, GSa), desc(128, GSw), idesc(n, 0, 1), K // 8, 256, GSw, n)
print("K-major A, MN-major B (interleave, swapped):", relerr(got, A.astype(np.float64) @ W.astype(np.float64)))
