"""steps/s of hidden-layer packs of n models: FFMA cluster kernel (auto cluster) vs tensor-core kernel (auto cluster)."""
import sys, os, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
def run(wl, n_models, K, tc):
    os.environ["VAE_MLP_TC"] = tc
    os.environ.pop("VAE_MLP_TC_CLUSTER", None)
    specs = bench.make_specs(wl, n_models)
    runs = PackedRuns(specs, max_chunk=K)
    runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(K); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); runs.train_steps(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{wl} models={n_models:4d} tc={tc} cluster {runs.plan.cluster:2d}: {ms/K*1e3:7.1f} us/step {n_models*K/ms*1e3:9.0f} steps/s", flush=True)
    del runs
    torch.cuda.empty_cache()
for nm in (2, 8, 15, 16, 18, 19, 24, 33, 37, 48, 64, 74, 100, 148, 296):
    for tc in ("0", "1"):
        run("sphere_dd3_pd3_ld6", nm, 100, tc)
