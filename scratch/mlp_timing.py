"""Cycle breakdown of one step of the FFMA cluster kernel (overlapped schedule), as thread 0 of the first and the last
CTA of cluster 0 see it.  Needs a -DVAE_MLP_TIMING build:  VAE_B200_DEFINES=VAE_MLP_TIMING VAE_B200_LIB=.../libvae_b200_timing.so"""
import sys, os, ctypes as C, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
os.environ["VAE_MLP_TC"] = "0"
lib = C.CDLL(os.environ["VAE_B200_LIB"])
names = ["top", "R+xform", "fwd load wait", "fwd gemm", "fwd barrier", "output/loss", "bwd first load", "bwd dX||dW tiles",
         "bwd barrier", "bwd issue loads", "bwd adam", "bwd load wait", "latent"]
order = [12, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]     # enum order in fused_mlp.cu: kTR.. ; kTTop is last
for nm in (1, 2):
    specs = bench.make_specs("sphere_dd3_pd3_ld6", nm)
    runs = PackedRuns(specs, max_chunk=500)
    runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(200); torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)()
    lib.vae_mlp_timing(out, 1)
    K = 500
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); runs.train_steps(K); e1.record(); torch.cuda.synchronize()
    lib.vae_mlp_timing(out, 1)
    print(f"models={nm} cluster {runs.plan.cluster}: {e0.elapsed_time(e1) / K * 1e3:.1f} us/step")
    for cta in (0, 1):
        v = [out[cta * 16 + i] / K for i in range(13)]
        tot = sum(v)
        print(f"  CTA {'first' if cta == 0 else 'last'}: total {tot:.0f} cycles/step")
        # enum: kTR=0, kTFwdLoad, kTFwdGemm, kTFwdBar, kTOut, kTBwdFirstLoad, kTBwdTiles, kTBwdBar, kTBwdIssue, kTBwdAdam, kTBwdWait, kTLatent, kTTop
        labels = ["R+xform", "fwd load wait", "fwd gemm", "fwd barrier", "output/loss", "bwd first load", "bwd dX||dW tiles",
                  "bwd barrier", "bwd issue loads", "bwd adam", "bwd load wait", "latent", "top"]
        for lab, x in zip(labels, v):
            print(f"    {lab:18s} {x:9.0f}  {100 * x / tot:5.1f}%")
