"""us/step of the hidden-layer kernels for small packs: FFMA cluster kernel vs tensor-core kernel at each cluster size."""
import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
def run(wl, n_models, K, tc, cs):
    os.environ["VAE_MLP_TC"] = tc
    if cs: os.environ["VAE_MLP_TC_CLUSTER"] = cs
    else: os.environ.pop("VAE_MLP_TC_CLUSTER", None)
    specs = bench.make_specs(wl, n_models)
    runs = PackedRuns(specs, max_chunk=K)
    runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(K); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); runs.train_steps(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{wl} tc={tc} cs={cs} models={n_models} K={K}: {ms/K*1e3:.1f} us/step, {n_models*K/ms*1e3:.0f} steps/s, plan cluster {runs.plan.cluster} tc {runs.plan.tc}, loss {float(runs.loss[K-1]):.5f}", flush=True)
models = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 15]
for wl in ("sphere_dd3_pd3_ld6", "sphere_dd5_pd16_ld16"):
    for nm in models:
        run(wl, nm, 200, "0", None)
        for cs in ("4", "8", "16"):
            try:
                run(wl, nm, 200, "1", cs)
            except Exception as e:
                print("FAILED", wl, nm, cs, repr(e)[:300], flush=True)
    if len(sys.argv) > 2: break
