import sys, torch
sys.path.insert(0, '.')
from vae_training_b200 import ops
def bench(M,N,K, iters=5):
    A = torch.randn(M,K,device='cuda'); B = torch.randn(N,K,device='cuda')
    Ah,Al = ops.split_tf32(A); Bh,Bl = ops.split_tf32(B)
    C = torch.empty(M,N,device='cuda')
    for _ in range(2): ops.gemm_nt_3xtf32(Ah,Al,Bh,Bl,out=C)
    torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm_nt_3xtf32(Ah,Al,Bh,Bl,out=C)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/iters
    # fp32 torch (cuBLAS SGEMM) and tf32 for reference
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(2): torch.matmul(A, B.t(), out=C)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): torch.matmul(A, B.t(), out=C)
    e1.record(); torch.cuda.synchronize(); ms32 = e0.elapsed_time(e1)/iters
    torch.backends.cuda.matmul.allow_tf32 = True
    for _ in range(2): torch.matmul(A, B.t(), out=C)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): torch.matmul(A, B.t(), out=C)
    e1.record(); torch.cuda.synchronize(); mstf = e0.elapsed_time(e1)/iters
    fl = 2.0*M*N*K
    print(f"{M}x{N}x{K}: ours {ms:.3f} ms = {fl/ms/1e9:.1f} TFLOP/s (algorithmic) | cuBLAS fp32 {ms32:.3f} ms = {fl/ms32/1e9:.1f} | cuBLAS tf32 {mstf:.3f} ms = {fl/mstf/1e9:.1f}")
bench(8192,4096,4096)
bench(65536,4096,4096, iters=3)
bench(65536,4096,1024, iters=3)
bench(4096,4096,65536, iters=3)
