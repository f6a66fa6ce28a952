"""cta_group::2 GEMM (VAE_TC_2SM=1) vs the multicast pair kernel: correctness vs float64 and timing."""
import os, sys, torch
sys.path.insert(0, '.')
from vae_training_b200 import ops
def run(M, N, K, iters=3):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda')
    bias = torch.randn(N, device='cuda')
    Ah, Al = ops.split_tf32(A); Bh, Bl = ops.split_tf32(B)
    C = torch.empty(M, N, device='cuda')
    idx = torch.cat([torch.arange(0, min(M, 256)), torch.arange(max(0, M - 200), M)]).cuda()
    ref = torch.relu(A[idx].double() @ B.double().t() + bias.double())
    for mode in ("0", "1"):
        os.environ["VAE_TC_2SM"] = mode
        for _ in range(2): ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, bias=bias, relu=True, out=C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, bias=bias, relu=True, out=C)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        err = float((C[idx].double() - ref).abs().max() / ref.abs().max())
        print(f"{M}x{N}x{K} 2sm={mode}: {ms:.3f} ms = {2.0*M*N*K/ms/1e9:.1f} TFLOP/s algorithmic, rel err {err:.2e}", flush=True)
run(256, 128, 64, 1)
run(1152, 256, 512)
run(4097, 520, 96)
run(8192, 4096, 4096)
run(65536, 4096, 1024)
run(65536, 4096, 4096)
