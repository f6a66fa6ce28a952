import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns, OverlappedPack
K = 20
specs = bench.make_specs("sigmoid_dd3_pd3_ld6", 64)
states = [init.initial_state(s) for s in specs]
runs = PackedRuns(specs, max_chunk=K)
runs.upload(states); runs.prepare(); runs.train_steps(K); torch.cuda.synchronize()
ref = runs.loss.view(64, K).cpu().numpy().copy()
for parts in (1, 4):
    op = OverlappedPack(specs, max_chunk=K, parts=parts)
    op.stage(states); op.prepare()
    for it in range(4):
        op.train_host_all(K, graph=True)
        torch.cuda.synchronize()
        got = op.host_losses(K)
        print("parts", parts, "call", it, "graph" if it >= 1 else "eager", "equal", np.array_equal(got, ref), "maxdiff", float(np.abs(got - ref).max()), "got[0,:3]", got[0, :3], "ref[0,:3]", ref[0, :3], flush=True)
        hk = op.packs[0].hkeys.numpy()[:8]
        print("   host keys after:", hk, "step", op.packs[0].hstep.numpy()[:1])
