"""A/B of the FFMA cluster kernel: us/step for 1, 2, 15 models (lib chosen by VAE_B200_LIB)."""
import sys, os, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
os.environ["VAE_MLP_TC"] = "0"
for wl in ("sphere_dd3_pd3_ld6", "sphere_dd5_pd16_ld16"):
    for nm in (1, 2, 15):
        specs = bench.make_specs(wl, nm)
        runs = PackedRuns(specs, max_chunk=500)
        runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(500); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); runs.train_steps(500); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{os.path.basename(os.environ.get('VAE_B200_LIB','default'))} {wl} models={nm} cluster {runs.plan.cluster}: {best/500*1e3:.1f} us/step loss {float(runs.loss[499]):.6f}", flush=True)
