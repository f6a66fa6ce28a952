"""One 16384 x 4096 x 4096 3xTF32 GEMM (for ncu captures of the tcgen05 kernel)."""
import sys, torch
sys.path.insert(0, '.')
from vae_training_b200 import ops
M, N, K = 16384, 4096, 4096
A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda')
Ah, Al = ops.split_tf32(A); Bh, Bl = ops.split_tf32(B)
C = torch.empty(M, N, device='cuda')
for _ in range(3):
    ops.gemm_nt_3xtf32(Ah, Al, Bh, Bl, out=C)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
