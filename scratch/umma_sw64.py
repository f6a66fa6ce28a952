import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_umma_probe import trunc_tf32, run, relerr, idesc
from vae_training_b200 import _lib as lib
def sw64_image(mat):
    rows, cols = mat.shape
    assert cols == 16 and rows % 8 == 0
    img = np.zeros((rows, 4, 4), np.float32)
    blk = mat.reshape(rows, 4, 4)
    for r in range(rows):
        perm = np.arange(4) ^ ((r >> 1) & 3)
        img[r][perm] = blk[r]
    return img.reshape(-1)
def desc(lbo, sbo, lt):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (lt << 61)
rng = np.random.default_rng(5)
for n in (104, 208):
    A = trunc_tf32(rng.standard_normal((128, 16)).astype(np.float32)); B = trunc_tf32(rng.standard_normal((n, 16)).astype(np.float32))
    got = run(lib, sw64_image(A), sw64_image(B), desc(16, 512, 4), desc(16, 512, 4), idesc(n), 2, 32, 32, n)
    print("SW64 K-major N", n, relerr(got, A.astype(np.float64) @ B.astype(np.float64).T))
    cyc = run(lib, sw64_image(A), sw64_image(B), desc(16, 512, 4), desc(16, 512, 4), idesc(n), 2, 32, 32, n, repeat=128, cycles=True)
    print("   cycles/MMA", cyc / 256)
