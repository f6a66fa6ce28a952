"""Race stress of the tensor-core kernel's cluster mode: 2000 steps at every cluster size must reproduce the 1-CTA run bit for bit."""
import sys, os, torch
sys.path.insert(0, '/root/repo')
import bench
from vae_training_b200 import init
from vae_training_b200.packed import PackedRuns
os.environ["VAE_MLP_TC"] = "1"
N = 2000
def run(wl, nm, cs):
    os.environ["VAE_MLP_TC_CLUSTER"] = cs
    specs = bench.make_specs(wl, nm)
    runs = PackedRuns(specs, max_chunk=N)
    runs.upload([init.initial_state(s) for s in specs]); runs.prepare()
    assert runs.plan.tc == 1 and runs.plan.cluster == int(cs)
    for chunk in (700, 1300):
        runs.train_steps(chunk)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(runs.loss.view(nm, N)).all())
    return runs.arena.view(torch.int32).clone()      # keys / step counters live in the arena too: compare bit patterns
ok = True
for wl, nm in (("sphere_dd3_pd3_ld6", 18), ("sphere_dd5_pd16_ld16", 7)):
    ref = run(wl, nm, "1")
    for cs in ("2", "4", "8"):
        for rep in range(2):
            got = run(wl, nm, cs)
            same = torch.equal(ref, got)
            ok &= same
            print(wl, nm, "cs", cs, "rep", rep, "bit-identical" if same else f"MISMATCH {int((ref != got).sum())} words", flush=True)
print("STRESS OK" if ok else "STRESS FAILED")
