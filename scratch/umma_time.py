import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_umma_probe import sw128_image, run, desc_hi, idesc
from vae_training_b200 import _lib as lib
def img_core(mat):
    rows, cols = mat.shape
    return np.ascontiguousarray(mat.reshape(rows // 8, 8, cols // 4, 4).transpose(0, 2, 1, 3)).reshape(-1).astype(np.float32)
rng = np.random.default_rng(0)
for n in (104, 200, 208, 256):
    A = rng.standard_normal((128, 32)).astype(np.float32); B = rng.standard_normal((n, 32)).astype(np.float32)
    for rep in (1, 64):
        c_sw = run(lib, sw128_image(A), sw128_image(B), desc_hi(16, 1024), desc_hi(16, 1024), idesc(n), 4, 32, 32, n, repeat=rep, cycles=True)
        K = 32; GS = K * 32
        d0 = ((128 >> 4) << 16) | ((GS >> 4) << 32) | (1 << 46)
        c_ns = run(lib, img_core(A), img_core(B), d0, d0, idesc(n), 4, 256, 256, n, repeat=rep, cycles=True)
        print(f"N={n} MMAs={4*rep}: SW128 {c_sw:.0f} cycles ({c_sw/(4*rep):.0f}/MMA), no-swizzle {c_ns:.0f} cycles ({c_ns/(4*rep):.0f}/MMA)")
