import sys, os, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_parity as T
from vae_training_b200 import ops, _lib
case_ids = [int(a) for a in sys.argv[1:]] or [5, 6, 0]
n = 12
for ci in case_ids:
    case = T.MLP_CASES[ci]
    res = {}
    for tc in ("1", "0"):
        os.environ["VAE_MLP_TC"] = tc
        d, ds, tr, st, job = T.mlp_setup(ops, _lib, case, n)
        ops.train_fused([job], n, ops.fused_scratch(1))
        torch.cuda.synchronize()
        res[tc] = (st["loss"].cpu().numpy().astype(np.float64), st["P"].cpu().numpy().copy(), st["mk"].cpu().numpy(), st["dk"].cpu().numpy())
    tr.train(10 ** 9, start=0, stop=n, events=False)
    ref = np.array(tr.train_losses, np.float64)
    for tc in ("1", "0"):
        got, P, mk, dk = res[tc]
        print(case[:6], "TC" if tc == "1" else "FFMA", "loss rel", np.max(np.abs(got - ref) / np.abs(ref)), "keys", (mk == tr.key).all(), (dk == ds.key).all(),
              "param maxabs", np.abs(P - tr.params).max(), "bad frac", np.mean(np.abs(P - tr.params) > n * 1e-3 * d.lr + 1e-7))
    print("  first losses TC", res["1"][0][:4], "ref", ref[:4])
