import sys, os, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init, ops
from vae_training_b200.packed import PackedRuns
def run(n_models, K, tc):
    os.environ["VAE_MLP_TC"] = tc
    specs = bench.make_specs("sphere_dd3_pd3_ld6", n_models)
    runs = PackedRuns(specs, max_chunk=K)
    runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(K); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); runs.train_steps(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"tc={tc} models={n_models} K={K}: {ms/K*1e3:.1f} us/step, {n_models*K/ms*1e3:.0f} steps/s, plan cluster {runs.plan.cluster}, loss {float(runs.loss[K-1]):.4f}", flush=True)
for nm in (1, 15, 148):
    for tc in ("1", "0"):
        run(nm, 200, tc)
