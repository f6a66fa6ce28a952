import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init, ops, _lib
from vae_training_b200.packed import PackedRuns
lib = C.CDLL(os.environ["VAE_B200_LIB"])
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
K = 100
specs = bench.make_specs("sphere_dd3_pd3_ld6", nm)
runs = PackedRuns(specs, max_chunk=K)
runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(K); torch.cuda.synchronize()
out = (C.c_longlong * 16)()
lib.vae_tc_profile(out, 1)
runs.train_steps(K); torch.cuda.synchronize()
lib.vae_tc_profile(out, 0)
names = ["gemm chunk loop", "gemm + final acc wait", "adam normal", "adam transposed", "fwd epilogue", "dX epilogue", "rng + data", "whole step",
         "  producer: empty wait", "  producer: tma issue", "  mma thread: ready wait", "  mma thread: issue + commit",
         "  converter: full wait", "  converter: convert + arrive"]
for n, v in zip(names, list(out)[:14]): print(f"{n:24s} {v / K:12.0f} cycles/step")
