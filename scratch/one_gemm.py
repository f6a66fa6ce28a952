import sys, torch
sys.path.insert(0, '.')
from vae_training_b200 import ops
M,N,K = 16384,4096,4096
A = torch.randn(M,K,device='cuda'); B = torch.randn(N,K,device='cuda')
Ah,Al = ops.split_tf32(A); Bh,Bl = ops.split_tf32(B)
C = torch.empty(M,N,device='cuda')
for _ in range(3): ops.gemm_nt_3xtf32(Ah,Al,Bh,Bl,out=C)
torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.gemm_nt_3xtf32(Ah,Al,Bh,Bl,out=C); e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1), "TFLOP/s", 2.0*M*N*K/e0.elapsed_time(e1)/1e9)
