"""A/B of the GEMM cluster variants (VAE_TC_CLUSTER=1|2|4), correctness vs float64 on a slice."""
import os, sys, torch
sys.path.insert(0, '.')
from vae_training_b200 import ops
def run(M, N, K, iters=3):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda')
    Ah, Al = ops.split_tf32(A); Bh, Bl = ops.split_tf32(B)
    C = torch.empty(M, N, device='cuda')
    ref = (A[:256].double() @ B.double().t())
    for cap in ("1", "2", "4"):
        os.environ["VAE_TC_CLUSTER"] = cap
        for _ in range(2): ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, out=C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, out=C)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        err = float((C[:256].double() - ref).abs().max() / ref.abs().max())
        err2 = float((C[-256:].double() - (A[-256:].double() @ B.double().t())).abs().max() / ref.abs().max())
        print(f"{M}x{N}x{K} cluster {cap}: {ms:.3f} ms = {2.0*M*N*K/ms/1e9:.1f} TFLOP/s algorithmic, rel err {err:.2e} {err2:.2e}", flush=True)
run(8192, 4096, 4096)
run(65536, 4096, 4096)
run(65536, 4096, 1024)
run(65536, 1024, 4096)
run(4097, 4096, 65536)
run(1024 + 128, 256, 512)
