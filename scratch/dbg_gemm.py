import sys, torch
sys.path.insert(0, '.')
from vae_training_b200 import ops
torch.manual_seed(0)
for (M,N,K) in [(128,128,256),(128,128,4096),(256,256,512)]:
    A = torch.randn(M,K,device='cuda'); B = torch.randn(N,K,device='cuda')
    ref = A.double() @ B.double().t(); scale = float(ref.abs().max())
    def trunc(x):
        h = (x.view(torch.int32) & -8192).view(torch.float32); return h, x - h
    for name, (Ah,Al),(Bh,Bl) in [("round", ops.split_tf32(A), ops.split_tf32(B)), ("trunc", trunc(A), trunc(B))]:
        outs = [ops.gemm_nt_3xtf32(Ah,Al,Bh,Bl).clone() for _ in range(3)]
        torch.cuda.synchronize()
        det = all(torch.equal(outs[0], o) for o in outs)
        e = (outs[0].double()-ref).abs()
        # emulate: exact products of splits in fp64
        emu = Ah.double()@Bh.double().t() + Ah.double()@Bl.double().t() + Al.double()@Bh.double().t()
        e_emu = float((outs[0].double()-emu).abs().max())/scale
        e_split = float((emu-ref).abs().max())/scale
        hh = Ah.double()@Bh.double().t()
        e_hh = float((outs[0].double()-hh).abs().max())/scale
        print(M,N,K,name,"det",det,"err",float(e.max())/scale,"vs_emulated",e_emu,"split_err",e_split,"vs_hi_only",e_hh, "rowmax argmax", int(e.max(dim=1).values.argmax()), "colmax argmax", int(e.max(dim=0).values.argmax()))
    e32 = float(((A@B.t()).double()-ref).abs().max())/scale
    print("   fp32 torch err", e32)
