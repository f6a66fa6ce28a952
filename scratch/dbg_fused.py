import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import datasets as ods, trainer as otr, vae_step as vs
from vae_training_b200 import _lib, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ds = ods.SigmoidDataset(69, 3, 3)
d = vs.VaeDesc(100, 7, 6, sig_decoder=True, epsilon0=-3.0, lr=1e-4)
flat = vs.init_params(d, 0)
tr = otr.OracleTrainer(d, ds, flat, dtype=np.float64); tr.begin()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
P, M, V = dev(flat), torch.zeros(d.n_params, device="cuda"), torch.zeros(d.n_params, device="cuda")
step = torch.zeros(1, dtype=torch.int32, device="cuda")
mk, dk = ops.key_tensor(tr.key), ops.key_tensor(ds.key)
loss = torch.zeros(n, device="cuda")
cdesc = _lib.make_desc(100, 7, 6, sig_decoder=True, tunable_decoder_var=True, epsilon0=-3.0, lr=1e-4)
job = ops.make_job(cdesc, _lib.make_data_desc("sigmoid", 3, 3), P, M, V, step, mk, dk, dev(ds.A), loss)
ops.train_fused([job], n, ops.fused_scratch(1)); torch.cuda.synchronize()
tr.train(10**9, 0, n, events=False)
print("loss gpu", loss.cpu().numpy()[:8]); print("loss ref", np.array(tr.train_losses)[:8])
g = P.cpu().numpy(); 
for name, shape, off in d.segments():
    k = int(np.prod(shape)); print(name, np.abs(g[off:off+k]-tr.params[off:off+k]).max(), np.abs(M.cpu().numpy()[off:off+k]-tr.m[off:off+k]).max()/ (np.abs(tr.m[off:off+k]).max()+1e-30))
# ---- debug dump comparison
import os
if os.environ.get("DBG"):
    tr2 = otr.OracleTrainer(d, ods.SigmoidDataset(69,3,3), flat, dtype=np.float64); tr2.begin()
    from oracle import jax_random as jr
    ds2 = tr2.dataset
    x = ds2.get_batch(100)
    key, _ = jr.model_chain_next(tr2.key)
    z = jr.normal(key, (100, 13))
    dbg = torch.zeros(20000, device="cuda")
    P, M, V = dev(flat), torch.zeros(d.n_params, device="cuda"), torch.zeros(d.n_params, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    mk, dk = ops.key_tensor(tr2.key), ops.key_tensor(ods.SigmoidDataset(69,3,3).key)
    dsx = ods.SigmoidDataset(69,3,3); dsx.get_key(); dk = ops.key_tensor(dsx.key)
    job = ops.make_job(cdesc, _lib.make_data_desc("sigmoid", 3, 3), P, M, V, step, mk, dk, dev(ds.A), loss, workspace=dbg)
    ops.train_fused([job], 1, ops.fused_scratch(1)); torch.cuda.synchronize()
    h = dbg.cpu().numpy()
    total = 0
    # find offsets: stored after total; search
    for tot in range(1000, 19000):
        q = h[tot:tot+8].view(np.int32)
        if q[7] == 0 and 0 < q[0] < tot and q[0] < q[1] < q[2] < q[3] < q[4] < q[5] and q[6] < q[0]:
            total = tot; break
    q = h[total:total+8].view(np.int32); print("offsets", q, total)
    zz = h[q[0]:q[0]+1300].reshape(100,13); print("z diff", np.abs(zz - z).max())
    xx = h[q[2]:q[2]+500].reshape(100,5)[:, :4]; print("x diff", np.abs(xx - x[:, :4]).max())
    f = vs.forward(d, flat.astype(np.float64), x, z[:, :6], z[:, 6:], dtype=np.float64)
    zs = h[q[3]:q[3]+700].reshape(100,7)[:, :6]; print("zs diff", np.abs(zs - f["zs"]).max())
    r = f["xhat"] - x; dxh = r/(np.exp(f["e"])*100)
    dx = h[q[4]:q[4]+700].reshape(100,7); print("dxhat diff", np.abs(dx - dxh).max(), np.abs(dxh).max())
    print("red", h[q[6]:q[6]+12], "ref mu2", (f["mu"]**2).sum(), "r2", (r*r).sum(), "rz", (r*z[:,6:]).sum())
    print("x gpu", xx[:3]); print("x ref", x[:3, :4])
    gg = h[q[1]:q[1]+300].reshape(100,3); print("gn", gg[:3])
    bkey, _ = jr.dataset_chain_next(dsx.key); print("ref normal", jr.normal(bkey, (100,3))[:3])
    print("A", ds.A.ravel(), h[q[6]+16:q[6]+20])
