import sys, os, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
os.environ["VAE_MLP_TC"] = "1"
for wl in ("sphere_dd3_pd3_ld6", "sphere_dd5_pd16_ld16"):
    for nm, cs in ((148, "1"), (296, "1"), (148, "2"), (74, "2"), (66, "4"), (33, "4"), (1, "1"), (1, "8")):
        os.environ["VAE_MLP_TC_CLUSTER"] = cs
        specs = bench.make_specs(wl, nm)
        runs = PackedRuns(specs, max_chunk=100)
        runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(100); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); runs.train_steps(100); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{os.path.basename(os.environ.get('VAE_B200_LIB','default')):22s} {wl} models={nm:4d} cs={cs}: {best/100*1e3:7.1f} us/step {nm*100/best*1e3:9.0f} steps/s loss {float(runs.loss[99]):.5f}", flush=True)
        del runs; torch.cuda.empty_cache()
    break
