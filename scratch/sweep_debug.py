import sys, time, torch
sys.path.insert(0, '/root/repo')
from vae_training_b200 import sweep, run as runmod
from vae_training_b200.vae import train_many
runs = sweep.builtin_runs("all")
part = sweep.latency_partition(runs, 8)[7]
def go(idx, label):
    models = []
    for i in idx:
        a = runs[i]
        ds = runmod.get_dataset(a.dataset, a.dataset_seed, a.padding_dim, a.batch_size, a)
        models.append(runmod.get_model(a, ds, None, verbose=False))
    torch.cuda.synchronize()
    t = {}
    t0 = time.perf_counter()
    train_many(models, save=False, timings=t)
    torch.cuda.synchronize()
    print(label, [ (runs[i].dataset[:3], runs[i].dataset_dimension, runs[i].padding_dim) for i in idx], round(time.perf_counter() - t0, 2), {k: round(v, 2) for k, v in t.items()}, flush=True)
small = [i for i in part if runs[i].dataset != "sphere"]
sph = [i for i in part if runs[i].dataset == "sphere"]
go(small, "small only")
go(small[:1], "one small")
go(sph, "sphere only")
go(part, "both")
