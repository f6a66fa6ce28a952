"""Per-launch overhead of the headline pack: launch time vs steps per launch (linear fit)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
specs = bench.make_specs("sigmoid_dd3_pd3_ld6", 1184)
runs = PackedRuns(specs, max_chunk=200)
runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(50); torch.cuda.synchronize()
xs, ys = [], []
for K in (1, 2, 5, 10, 20, 50, 100, 200):
    R = 200
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    ev[0].record()
    for i in range(R):
        runs.train_steps(K); ev[i + 1].record()
    torch.cuda.synchronize()
    d = np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(R)]) * 1e3
    xs.append(K); ys.append(d)
    print(f"K={K:4d}: {d:8.2f} us per launch, {d / K:7.2f} us per step", flush=True)
a, b = np.polyfit(xs, ys, 1)
print(f"fit: {a:.3f} us per step + {b:.2f} us per launch")
