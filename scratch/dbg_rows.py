import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
os.environ.setdefault("PYTEST_CURRENT_TEST", "x")
from vae_training_b200 import ops, _lib
from oracle import datasets as ods, vae_step as vs, trainer as otr
import test_gpu_parity as T

def run(case, n, rows):
    os.environ["VAE_MLP_ROWS"] = rows
    d, ds, tr, st, job = T.mlp_setup(ops, _lib, case, n)
    ops.train_fused([job], n, ops.fused_scratch(1))
    torch.cuda.synchronize()
    return d, ds, tr, st

for case in [][:0] + [("sphere", 5, 16, 16, [200]*3, [200]*3, 100, 48), ("sphere", 5, 5, 10, [200]*3, [200]*3, 100, 69), ("sphere", 3, 3, 6, [200]*3, [200]*3, 100, 69),
             ("sphere", 5, 0, 4, [32], [32], 100, 69), ("sphere", 5, 3, 4, [], [32], 100, 1), ("sphere", 4, 0, 4, [32], [32], 16, 69)][:6]:
    n = 1
    try:
        d, ds, tr, a = run(case, n, "0")
        _, _, _, b = run(case, n, "1")
    except Exception as e:
        print(case, "ERR", e); continue
    tr.train(10 ** 9, start=0, stop=n, events=False)
    ref = np.array(tr.train_losses)
    la, lb = a["loss"].cpu().numpy()[:n].astype(np.float64), b["loss"].cpu().numpy()[:n].astype(np.float64)
    print(case[:4], "loss ref", ref, "col err", (la - ref) / ref, "rows err", (lb - ref) / ref)
    Pa, Pb = a["P"].cpu().numpy().astype(np.float64), b["P"].cpu().numpy().astype(np.float64)
    for nm, shape, off in d.segments():
        cnt = int(np.prod(shape))
        da = np.max(np.abs(Pa[off:off+cnt] - tr.params[off:off+cnt])); db = np.max(np.abs(Pb[off:off+cnt] - tr.params[off:off+cnt]))
        if db > 3 * da + 1e-7: print("   ", nm, shape, "col", da, "rows", db)

print("---- partial sums")
def ws_offsets(d, dd, G):
    up4 = lambda n: (n + 3) & ~3
    g = (d.batch + 3) // 4; g |= 1; BP = 4 * g
    off = 0
    def take(n):
        nonlocal off
        r = off; off += up4(n); return r
    es, dsz = d.enc_sizes, d.dec_sizes
    o = {}
    o["zT"] = take((d.latent_dim + d.data_dim) * BP); o["gn"] = take(d.batch * G)
    o["enc"] = [take(es[i + 1] * BP) for i in range(len(es) - 1)]
    o["dec"] = [take(dsz[i + 1] * BP) for i in range(len(dsz) - 1)]
    mw = max(es + dsz)
    o["g0"] = take(mw * BP); o["g1"] = take(mw * BP); o["dz"] = take(d.latent_dim * BP)
    o["xT"] = take(dd * BP); o["zsT"] = take(d.latent_dim * BP)
    o["dye"] = [take(es[i + 1] * BP) for i in range(len(es) - 1)]
    o["dyd"] = [take(dsz[i + 1] * BP) for i in range(len(dsz) - 1)]
    o["wte"] = [take(es[i] * es[i + 1]) for i in range(len(es) - 1)]
    o["wtd"] = [take(dsz[i] * dsz[i + 1]) for i in range(len(dsz) - 1)]
    o["part"] = take(16 * 68)
    o["BP"] = BP
    return o

case = ("sphere", 5, 16, 16, [200]*3, [200]*3, 100, 48)
d, ds, tr, b = run(case, 1, "1")
W = b["ws"].view(torch.float32).cpu().numpy().astype(np.float64)
o = ws_offsets(d, 5, 5)
BP = o["BP"]
part = W[o["part"]:o["part"] + 16 * 68].reshape(16, 68)
print("kernel sums r2, rz, mu2:", part[:, 0].sum(), part[:, 1].sum(), part[:, 2].sum())
# oracle forward on the same step
import oracle.jax_random as jr
tr2 = otr.OracleTrainer(d, ods.make_dataset(case[0], case[7], case[1], case[2]), vs.init_params(d, seed=0), dtype=np.float64)
tr2.begin()
x = tr2.dataset.get_batch(d.batch)
key, _ = jr.model_chain_next(tr2.key)
z = tr2.sample_latent(key, d.batch)
L = d.latent_dim
fo = vs.forward(d, tr2.params, x, z[:, :L], z[:, L:], dtype=np.float64)
xhat, mu, lv, e = fo['xhat'], fo['mu'], fo['lv'], fo['e']
r = xhat - x
print("oracle sums r2, rz, mu2:", (r * r).sum(), (r * z[:, L:]).sum(), (mu * mu).sum())
xT = W[o["xT"]:o["xT"] + 5 * BP].reshape(5, BP)[:, :100].T
print("x max err", np.abs(xT - x[:, :5]).max())
zs = W[o["zsT"]:o["zsT"] + L * BP].reshape(L, BP)[:, :100].T
s = np.exp(0.5 * lv)
print("zs max err", np.abs(zs - (mu + s * z[:, :L])).max(), "zs scale", np.abs(zs).max())
h = W[o["enc"][0]:o["enc"][0] + 200 * BP].reshape(200, BP)[:, :100].T
p = vs.unflatten(d, tr2.params)
h_ref = np.maximum(x @ p["Encoder/FC0/kernel"] + p["Encoder/FC0/bias"], 0)
print("enc h0 max err", np.abs(h - h_ref).max(), np.abs(h_ref).max())
h1 = W[o["enc"][1]:o["enc"][1] + 200 * BP].reshape(200, BP)[:, :100].T
h1_ref = np.maximum(h_ref @ p["Encoder/FC1/kernel"] + p["Encoder/FC1/bias"], 0)
print("enc h1 max err", np.abs(h1 - h1_ref).max(), np.abs(h1_ref).max(), "mean signed", (h1 - h1_ref).mean())
dy = W[o["dyd"][3]:o["dyd"][3] + d.data_dim * BP].reshape(d.data_dim, BP)[:, :100].T
print("dxhat max err", np.abs(dy - r * np.exp(-e) / 100).max(), np.abs(dy).max())
