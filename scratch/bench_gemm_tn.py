"""dW-shaped GEMM: NT on transposed splits (incl. the two transpose_split passes) vs TN straight from the batch-major splits."""
import sys, torch
sys.path.insert(0, '/root/repo')
from vae_training_b200 import ops
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rows, fin, fout in ((65536, 4096, 4096), (65536, 1024, 4096), (8192, 4096, 4096)):
    X = torch.randn(rows, fin, device="cuda"); dY = torch.randn(rows, fout, device="cuda")
    xh, xl = ops.split_tf32(X); dh, dl = ops.split_tf32(dY)
    out = torch.empty(fin, fout, device="cuda")
    flop = 2.0 * rows * fin * fout
    ms_tn = t(lambda: ops.gemm_tn_3xtf32(xh, xl, dh, dl, out=out))
    ref = out.clone()
    xth, xtl = ops.transpose_split_tf32(X); dth, dtl = ops.transpose_split_tf32(dY)
    ms_nt = t(lambda: ops.gemm_nt_3xtf32(xth, xtl, dth, dtl, out=out))
    ms_tr = t(lambda: (ops.transpose_split_tf32(X), ops.transpose_split_tf32(dY)))
    print(f"rows {rows} {fin}x{fout}: TN {ms_tn:.3f} ms ({flop/ms_tn/1e9:.0f} TFLOP/s) | NT {ms_nt:.3f} ms ({flop/ms_nt/1e9:.0f}) + transposes {ms_tr:.3f} ms | max|TN-NT| {float((ref-out).abs().max()):.3e} of {float(out.abs().max()):.1f}", flush=True)
    del X, dY, xh, xl, dh, dl, xth, xtl, dth, dtl
    torch.cuda.empty_cache()
