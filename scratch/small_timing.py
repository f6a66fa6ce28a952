import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
lib = C.CDLL(os.environ["VAE_B200_LIB"])
specs = bench.make_specs("sigmoid_dd3_pd3_ld6", 1184)
runs = PackedRuns(specs, max_chunk=200)
runs.upload([init.initial_state(s) for s in specs]); runs.prepare(); runs.train_steps(50); torch.cuda.synchronize()
for K in (1, 20):
    rows = []
    for rep in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        runs.train_steps(K); torch.cuda.synchronize()
        e0.record(); runs.train_steps(K); e1.record(); torch.cuda.synchronize()
        out = (C.c_ulonglong * 8)()
        lib.vae_small_timing(out)
        t = np.array(list(out)[:6], dtype=np.float64)
        rows.append(np.concatenate([(t[1:] - t[:-1]) / 1e3, [e0.elapsed_time(e1) * 1e3]]))
    m = np.median(np.array(rows), axis=0)
    print(f"K={K}: entry->init {m[0]:.2f} us | zero-fill {m[1]:.2f} | state load + keys + gmask {m[2]:.2f} | step loop {m[3]:.2f} | store {m[4]:.2f} | event-to-event {m[5]:.2f} us")
