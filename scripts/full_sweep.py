"""BASELINE.json configs[3]: the full seed x padding x latent sweep of the three reference
scripts (54 runs, 7.05 M train steps), packed on the GPU(s) with the reference's own event
schedule (stats every 5000 steps, plot/save events at 0, 50k, 100k, last).

    python scripts/full_sweep.py [--which all] [--num_batches N] [--out gpurun_out/sweep.json]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/full_sweep.py   # LPT partition, no communication
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_training_b200 import sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="all")
    ap.add_argument("--num_batches", type=int, default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--emulate", default=None, help="RANK/WORLD: run that rank's share of a WORLD-GPU sweep alone on this GPU")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    emulate = args.emulate is not None
    if emulate:
        rank, world = (int(v) for v in args.emulate.split("/"))
    if world > 1 and not emulate:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    runs = sweep.builtin_runs(args.which, args.num_batches)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    timings = {}
    models = sweep.run_sweep(runs, rank, world, save=False, verbose=False, timings=timings)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = sum(m.num_batches for m in models)
    summary = {"rank": rank, "world": world, "runs": len(models), "train_steps": steps, "seconds": dt,
               "steps_per_sec": steps / dt, **timings, "results": []}
    for m in models:
        losses = m.vae_losses
        summary["results"].append({
            "dataset": m.spec.dataset, "dd": m.spec.dataset_dimension, "pd": m.spec.padding_dim,
            "latent": m.spec.latent_dimension, "num_batches": m.num_batches, "fused_class": m.fused_class,
            "first_loss": float(losses[1]), "last_loss": float(losses[-1]), "finite": bool(np.isfinite(losses).all()),
            "final_encoder_logvar": [float(v) for v in m.var_enc[-1].cpu().numpy()],
            "final_decoder_logvar": float(m.var_dec[-1].cpu().numpy()[0]),
            "stats_last": {k: float(v[-1]) for k, v in m.stats.items()}})
    if world > 1 and not emulate:
        gathered = [None] * world
        dist.all_gather_object(gathered, summary)
        if rank == 0:
            tot = sum(g["train_steps"] for g in gathered)
            wall = max(g["seconds"] for g in gathered)
            print(json.dumps({"world": world, "runs": sum(g["runs"] for g in gathered), "train_steps": tot,
                              "seconds_max_over_ranks": wall, "steps_per_sec": tot / wall,
                              "per_rank_seconds": [round(g["seconds"], 2) for g in gathered],
                              "per_rank_train_seconds": [round(g.get("train_seconds", 0.0), 2) for g in gathered]}), flush=True)
            summary = {"ranks": gathered}
        dist.destroy_process_group()
    else:
        print(json.dumps({k: v for k, v in summary.items() if k != "results"}), flush=True)
    if args.out and rank == 0:
        with open(args.out, "w") as f:
            json.dump(summary, f)


if __name__ == "__main__":
    main()
