"""Hardware self-check of the tcgen05 operand forms fused_mlp_tc.cu relies on (vae_umma_probe): K-major and
MN-major SWIZZLE_128B shared-memory descriptors for kind::tf32, N = 104, and the handling of the low 13
mantissa bits of raw FP32 operands.  The images are built here in NumPy with the layout formula the kernel
uses (canonical UMMA SWIZZLE_128B: 128-byte rows, 16-byte chunk index XOR (row % 8), 8-row groups of 1024 B)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sw128_image(mat):
    """mat[rows][cols] (rows % 8 == 0, cols % 32 == 0) -> float32 image: for each 32-column atom a block of
    rows x 128 B; element (r, c): chunk = (c % 32) // 4 stored at chunk ^ (r % 8)."""
    rows, cols = mat.shape
    assert rows % 8 == 0 and cols % 32 == 0
    img = np.zeros((cols // 32, rows, 32), np.float32)
    for a in range(cols // 32):
        blk = mat[:, 32 * a:32 * a + 32].reshape(rows, 8, 4)
        for r in range(rows):
            perm = np.arange(8) ^ (r % 8)
            img[a, r].reshape(8, 4)[perm] = blk[r]
    return img.reshape(-1)


def desc_hi(lbo_bytes, sbo_bytes):
    return ((lbo_bytes >> 4) << 16) | ((sbo_bytes >> 4) << 32) | (1 << 46) | (2 << 61)


def idesc(n, a_mn=False, b_mn=False):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((128 >> 4) << 24)


def trunc_tf32(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rna_tf32(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def run(lib, a_img, b_img, da, db, idesc_, ksteps, a_step, b_step, n, repeat=1, cycles=False):
    ta, tb = torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda()
    out = torch.zeros(128 * n + 8, device="cuda")
    rc = lib.load().vae_umma_probe(C.c_void_p(ta.data_ptr()), a_img.nbytes, C.c_void_p(tb.data_ptr()), b_img.nbytes,
                                   C.c_uint64(da), C.c_uint64(db), C.c_uint32(idesc_), ksteps, a_step, b_step, n, repeat,
                                   C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    lib.check(rc, "vae_umma_probe")
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    return float(res[128 * n]) if cycles else res[:128 * n].reshape(128, n)


@pytest.fixture(scope="module")
def lib():
    from vae_training_b200 import _lib
    return _lib


def relerr(got, ref):
    return float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))


def test_kmajor_operands_and_raw_fp32_truncation(lib):
    """Both operands K-major, N = 104, K = 32: also decides how raw FP32 operands are narrowed to TF32."""
    rng = np.random.default_rng(0)
    n = 104
    A = rng.standard_normal((128, 32)).astype(np.float32)
    B = rng.standard_normal((n, 32)).astype(np.float32)
    got = run(lib, sw128_image(A), sw128_image(B), desc_hi(16, 1024), desc_hi(16, 1024), idesc(n), 4, 32, 32, n)
    ref_t = trunc_tf32(A).astype(np.float64) @ trunc_tf32(B).astype(np.float64).T
    ref_r = rna_tf32(A).astype(np.float64) @ rna_tf32(B).astype(np.float64).T
    et, er = relerr(got, ref_t), relerr(got, ref_r)
    print(f"raw fp32 operands: error vs truncation model {et:.2e}, vs round-to-nearest model {er:.2e}")
    assert et <= 2e-6, (et, er)           # the tensor core ignores the low 13 mantissa bits
    assert er > 20 * et


def image_rows(mat, row_bytes, swz_shift, swz_mask):
    """K-major image with rows of `row_bytes` (16-byte chunk index XOR ((row >> swz_shift) & swz_mask)); matrices wider
    than one row are stored as consecutive [K block][row][row_bytes] blocks."""
    rows, cols = mat.shape
    per = row_bytes // 4
    img = np.zeros((cols // per, rows, per // 4, 4), np.float32)
    for a in range(cols // per):
        blk = mat[:, per * a:per * a + per].reshape(rows, per // 4, 4)
        for r in range(rows):
            perm = np.arange(per // 4) ^ ((r >> swz_shift) & swz_mask)
            img[a, r][perm] = blk[r]
    return img.reshape(-1)


def desc_k(sbo, layout):
    return (1 << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


@pytest.mark.parametrize("n", [104, 208])
def test_kmajor_swizzle_64b_rows(lib, n):
    """How fused_mlp_tc.cu stages ROW-MAJOR operands: TMA box {16 floats, rows} with SWIZZLE_64B -> 64-byte rows,
    descriptor layout 4, SBO 512; the second K = 8 step starts 32 bytes further."""
    rng = np.random.default_rng(5)
    A = trunc_tf32(rng.standard_normal((128, 16)).astype(np.float32))
    B = trunc_tf32(rng.standard_normal((n, 16)).astype(np.float32))
    got = run(lib, image_rows(A, 64, 1, 3), image_rows(B, 64, 1, 3), desc_k(512, 4), desc_k(512, 4), idesc(n), 2, 32, 32, n)
    assert relerr(got, A.astype(np.float64) @ B.astype(np.float64).T) <= 2e-6


@pytest.mark.parametrize("n", [104, 208])
def test_kmajor_swizzle_32b_groups(lib, n):
    """How it stages GROUP-MAJOR operands ([k / 8][row][8]): TMA box {8 floats, rows, 2 groups} with SWIZZLE_32B ->
    two [row][32-byte] blocks, descriptor layout 6, SBO 256; the second K = 8 step is the next block (32 x rows bytes)."""
    rng = np.random.default_rng(6)
    A = trunc_tf32(rng.standard_normal((128, 16)).astype(np.float32))
    B = trunc_tf32(rng.standard_normal((n, 16)).astype(np.float32))
    got = run(lib, image_rows(A, 32, 2, 1), image_rows(B, 32, 2, 1), desc_k(256, 6), desc_k(256, 6), idesc(n), 2, 128 * 32, n * 32, n)
    assert relerr(got, A.astype(np.float64) @ B.astype(np.float64).T) <= 2e-6


def test_mn_major_tf32_needs_the_32b_base_swizzle(lib):
    """Why the kernel keeps feature-major copies instead of reading X^T / dY^T MN-major from the batch-major arrays:
    for kind::tf32 an MN-major operand is only read correctly from the SWIZZLE_128B_BASE32B layout (layout 1: 4-row
    atoms, 32-byte chunk index XOR (row % 4), SBO 512); with the plain SWIZZLE_128B image the MMA returns zeros."""
    rng = np.random.default_rng(2)
    k, n = 32, 104
    A = trunc_tf32(rng.standard_normal((128, k)).astype(np.float32))
    W = trunc_tf32(rng.standard_normal((k, 128)).astype(np.float32))    # [k][n]: B read MN-major (rows = K)
    ref = A.astype(np.float64) @ W.astype(np.float64)[:, :n]
    ka = desc_hi(16, 1024)

    def base32_image(mat):
        rows, cols = mat.shape
        img = np.zeros((cols // 32, rows, 4, 8), np.float32)
        for a in range(cols // 32):
            blk = mat[:, 32 * a:32 * a + 32].reshape(rows, 4, 8)
            for r in range(rows):
                img[a, r][np.arange(4) ^ (r % 4)] = blk[r]
        return img.reshape(-1)
    mn = lambda lbo, sbo, layout: ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)
    good = run(lib, sw128_image(A), base32_image(W), ka, mn(k * 128, 512, 1), idesc(n, b_mn=True), 4, 32, 1024, n)
    assert relerr(good, ref) <= 2e-6
    bad = run(lib, sw128_image(A), sw128_image(W), ka, mn(k * 128, 1024, 2), idesc(n, b_mn=True), 4, 32, 1024, n)
    assert np.count_nonzero(bad) == 0


def base32_image(mat):
    """MN-major image for kind::tf32: mat[k][mn] (k rows, mn % 32 == 0) -> for each 32-wide mn block a [k][128 B] tile whose
    32-byte chunk index is XOR-ed with (k % 4)  (= TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B of a {32, k} box)."""
    rows, cols = mat.shape
    img = np.zeros((cols // 32, rows, 4, 8), np.float32)
    for a in range(cols // 32):
        blk = mat[:, 32 * a:32 * a + 32].reshape(rows, 4, 8)
        for r in range(rows):
            img[a, r][np.arange(4) ^ (r % 4)] = blk[r]
    return img.reshape(-1)


@pytest.mark.parametrize("n", [128, 256])
def test_mn_major_a_and_b_base32(lib, n):
    """The dW form of the large-model GEMM (gemm_tc.cu, TN mode): A = X^T and B = dY^T read MN-major straight from
    batch-major [k = batch row][feature] tiles -- both operands in the SWIZZLE_128B_BASE32B layout (descriptor layout 1,
    LBO = bytes between 32-wide feature blocks, SBO = 512, +1024 B per K = 8 step), instruction-descriptor bits 15 / 16
    set; N = 256 spans two 128-column operand tiles laid out back to back (the stacked [Bh ; Bl] form)."""
    rng = np.random.default_rng(11)
    k = 32
    X = trunc_tf32(rng.standard_normal((k, 128)).astype(np.float32))     # [k][m]
    Y = trunc_tf32(rng.standard_normal((k, n)).astype(np.float32))       # [k][n]
    mn = lambda lbo, sbo: ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (1 << 61)
    got = run(lib, base32_image(X), base32_image(Y), mn(k * 128, 512), mn(k * 128, 512), idesc(n, a_mn=True, b_mn=True),
              k // 8, 1024, 1024, n)
    ref = X.astype(np.float64).T @ Y.astype(np.float64)
    assert relerr(got, ref) <= 2e-6
