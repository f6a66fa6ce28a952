import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_umma_probe import trunc_tf32, run, relerr, idesc
from vae_training_b200 import _lib as lib
def sw32_image(mat):
    """[rows][16] as two K groups of 8: [g][row][32 B], 16-byte chunk c stored at c ^ ((row >> 2) & 1)"""
    rows, cols = mat.shape
    img = np.zeros((cols // 8, rows, 2, 4), np.float32)
    for g in range(cols // 8):
        blk = mat[:, 8 * g:8 * g + 8].reshape(rows, 2, 4)
        for r in range(rows):
            perm = np.arange(2) ^ ((r >> 2) & 1)
            img[g, r][perm] = blk[r]
    return img.reshape(-1)
def desc(lbo, sbo, lt):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (lt << 61)
rng = np.random.default_rng(5)
for n in (104, 208):
    A = trunc_tf32(rng.standard_normal((128, 16)).astype(np.float32)); B = trunc_tf32(rng.standard_normal((n, 16)).astype(np.float32))
    got = run(lib, sw32_image(A), sw32_image(B), desc(16, 256, 6), desc(16, 256, 6), idesc(n), 2, 128 * 32, n * 32, n)
    print("SW32 K-major N", n, relerr(got, A.astype(np.float64) @ B.astype(np.float64).T))
    cyc = run(lib, sw32_image(A), sw32_image(B), desc(16, 256, 6), desc(16, 256, 6), idesc(n), 2, 128 * 32, n * 32, n, repeat=128, cycles=True)
    print("   cycles/MMA", cyc / 256)
