"""Sweep launcher: parses the reference's ``*.sh`` experiment files (one ``python run.py ...``
per line) through the run.py flag surface and trains all runs packed on this rank's GPU,
or an LPT partition of them when launched under torchrun (one process per GPU, no
communication -- the runs are independent).

    python -m vae_training_b200.sweep seed_linpadding_expts.sh sigmoid_vae_padding_expts.sh ...
    python -m vae_training_b200.sweep builtin:all          # the 54 runs, regenerated from their grids
"""
from __future__ import annotations

import os
import shlex
import sys
from typing import List

from . import run as runmod


def parse_script(path: str) -> List:
    """Every `python run.py ...` line of a sweep script -> argparse namespace."""
    out = []
    with open(path) as f:
        for line in f:
            toks = shlex.split(line.strip())
            if len(toks) >= 3 and toks[0].startswith("python") and toks[1].endswith("run.py"):
                out.append(runmod.parse_arguments(toks[2:]))
    return out


def run_cost(args) -> float:
    """Relative cost of a run: steps x algorithmic flops per step."""
    from .packed import RunSpec
    from .vae import _parse_sizes
    s = RunSpec(dataset=args.dataset, dataset_dimension=args.dataset_dimension, padding_dim=args.padding_dim,
                latent_dimension=args.latent_dimension, layer_sizes=_parse_sizes(args.layer_sizes),
                encoder_layer_sizes=_parse_sizes(args.encoder_layer_sizes), batch_size=args.batch_size)
    return float(args.num_batches) * (s.flops_per_step() + 150.0 * s.normals_per_step())


def lpt_partition(costs: List[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first partition of run indices over n_parts GPUs."""
    parts = [[] for _ in range(n_parts)]
    load = [0.0] * n_parts
    for i in sorted(range(len(costs)), key=lambda i: -costs[i]):
        p = min(range(n_parts), key=lambda q: load[q])
        parts[p].append(i)
        load[p] += costs[i]
    return parts


def run_sweep(all_args, rank=0, world=1, save=True, verbose=False, timings=None):
    import time

    import torch

    from .vae import train_many
    t0 = time.perf_counter()
    mine = lpt_partition([run_cost(a) for a in all_args], world)[rank]
    models = []
    for i in mine:
        a = all_args[i]
        out_dir = runmod.make_output_dir(a.name, a.overwrite, a) if save else None
        ds = runmod.get_dataset(a.dataset, a.dataset_seed, a.padding_dim, a.batch_size, a)
        models.append(runmod.get_model(a, ds, out_dir, verbose=verbose))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    train_many(models, save=save)
    torch.cuda.synchronize()
    if timings is not None:
        timings["setup_seconds"] = t1 - t0                    # datasets, initial parameters, uploads
        timings["train_seconds"] = time.perf_counter() - t1   # the training loops incl. eval events
    if save:
        for m in models:
            m.save(final=True)
    return models


def builtin_runs(which: str, num_batches=None) -> List:
    """The reference's sweeps regenerated from their grids (sweeps.py); `which` in
    linear | sigmoid | sphere | all.  num_batches overrides the run length (smoke tests)."""
    from . import sweeps
    out = []
    for argv in sweeps.BUILTIN[which]():
        a = runmod.parse_arguments(argv)
        if num_batches is not None:
            a.num_batches = int(num_batches)
        out.append(a)
    return out


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    all_args = []
    num_batches = None
    if "--num_batches" in argv:
        i = argv.index("--num_batches")
        num_batches = int(argv[i + 1])
        argv = argv[:i] + argv[i + 2:]
    for path in argv:
        if path.startswith("builtin:"):
            all_args += builtin_runs(path.split(":", 1)[1], num_batches)
        else:
            all_args += parse_script(path)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if "LOCAL_RANK" in os.environ:
        import torch
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    models = run_sweep(all_args, rank, world, verbose=(rank == 0))
    print(f"[rank {rank}] trained {len(models)} of {len(all_args)} runs", flush=True)


if __name__ == "__main__":
    main()
