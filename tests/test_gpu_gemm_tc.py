"""tcgen05 3xTF32 GEMM building blocks (configs[4] path) against float64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vae_training_b200 import ops as _ops
    return _ops


def test_split_is_exact(ops):
    x = torch.randn(100003, device="cuda") * torch.logspace(-3, 3, 100003, device="cuda")
    x = x[: 100000].contiguous()
    hi, lo = ops.split_tf32(x)
    assert float(((hi.double() + lo.double() - x.double()).abs() / x.abs().double()).max()) <= 2.0 ** -22
    for t in (hi, lo):      # both parts are exactly representable in TF32
        assert torch.equal(t.view(torch.int32) & 0x1FFF, torch.zeros_like(t, dtype=torch.int32))
    assert float((lo.abs() / x.abs().clamp_min(1e-30)).max()) <= 2.0 ** -11 * 1.001
    a = torch.randn(37, 53, device="cuda")
    hT, lT = ops.transpose_split_tf32(a)
    assert hT.shape == (53, 37) and float((hT + lT - a.t()).abs().max()) <= 1e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 96), (300, 200, 100), (1024, 512, 4096),
                                   (100, 16, 4096), (4096, 1024, 16)])
def test_gemm_nt_3xtf32_matches_float64(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    Ah, Al = ops.split_tf32(A)
    Bh, Bl = ops.split_tf32(B)
    C = ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    scale = float(ref.abs().max())
    err = float((C.double() - ref).abs().max()) / scale
    err32 = float(((A @ B.t()).double() - ref).abs().max()) / scale      # plain FP32 for comparison
    assert err <= 1.5e-6, (err, err32)
    assert err <= 4 * err32 + 1e-7, (err, err32)


def test_gemm_epilogue_bias_relu_mask(ops):
    M, N, K = 200, 72, 64
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    bias, mask = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda")
    Ah, Al = ops.split_tf32(A)
    Bh, Bl = ops.split_tf32(B)
    ref = A.double() @ B.double().t()
    got = ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, bias=bias, relu=True)
    assert float((got.double() - torch.relu(ref + bias.double())).abs().max()) <= 1e-4
    got = ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, mask=mask)
    assert float((got.double() - ref * (mask > 0)).abs().max()) <= 1e-4


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 96), (300, 200, 100), (512, 1024, 4096),
                                   (4096, 256, 8192), (1028, 132, 1000)])
@pytest.mark.parametrize("cluster", ["2", "1"])
def test_gemm_tn_3xtf32_matches_float64(ops, M, N, K, cluster, monkeypatch):
    """The dW form: C = At^T Bt with both operands K-row-major, read MN-major by the tensor core from TMA
    SWIZZLE_128B_ATOM_32B boxes (ragged M / N / K: out-of-range columns and rows are zero-filled)."""
    monkeypatch.setenv("VAE_TC_CLUSTER", cluster)
    g = torch.Generator(device="cuda").manual_seed(M * 5 + N * 11 + K)
    At = torch.randn(K, M, device="cuda", generator=g)
    Bt = torch.randn(K, N, device="cuda", generator=g)
    Ah, Al = ops.split_tf32(At)
    Bh, Bl = ops.split_tf32(Bt)
    C = ops.gemm_tn_3xtf32(Ah, Al, Bh, Bl)
    torch.cuda.synchronize()
    ref = At.double().t() @ Bt.double()
    scale = float(ref.abs().max())
    err = float((C.double() - ref).abs().max()) / scale
    err32 = float(((At.t() @ Bt).double() - ref).abs().max()) / scale
    assert err <= 1.5e-6, (err, err32)
    assert err <= 4 * err32 + 1e-7, (err, err32)
