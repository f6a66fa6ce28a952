"""Sweep launcher: parses the reference's ``*.sh`` experiment files (one ``python run.py ...``
per line) through the run.py flag surface and trains all runs packed on this rank's GPU,
or a latency-aware partition of them (latency_partition) when launched under torchrun (one process per GPU, no
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


def _spec_of(args):
    from .packed import RunSpec
    from .vae import _parse_sizes
    return RunSpec(dataset=args.dataset, dataset_dimension=args.dataset_dimension, padding_dim=args.padding_dim,
                   latent_dimension=args.latent_dimension, layer_sizes=_parse_sizes(args.layer_sizes),
                   encoder_layer_sizes=_parse_sizes(args.encoder_layer_sizes), batch_size=args.batch_size)


def run_cost(args) -> float:
    """Relative cost of a run: steps x algorithmic flops per step."""
    s = _spec_of(args)
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


# Measured step latency (us) of a pack of n hidden-layer models on one B200 (profiles/r02_mlp_kernel_crossover.txt):
# every model of a resident pack advances one step in this time, so a rank's hidden-layer wall time is
# steps x chain_latency_us(n), not a sum of flops.  (n <= 2: 16-CTA FFMA clusters; <= 15: 8-CTA; 16 - 37: tensor-core
# kernel on 4-CTA clusters; 38 - 74: 2-CTA; up to 148: one CTA per model; beyond that the pack runs in waves.)
_CHAIN_LATENCY_US = ((2, 150.0), (15, 170.0), (37, 245.0), (74, 332.0), (148, 500.0))
_SMALL_STEP_US = 6.0          # a single-Dense chain in a resident pack (3.9 - 11 us per step by shape)


def chain_latency_us(n_models: int) -> float:
    if n_models <= 0:
        return 0.0
    for cap, us in _CHAIN_LATENCY_US:
        if n_models <= cap:
            return us
    return _CHAIN_LATENCY_US[-1][1] * ((n_models + 147) // 148)


def latency_partition(all_args, n_parts: int) -> List[List[int]]:
    """Partition of the runs over n_parts GPUs for WALL TIME.  The sweeps are latency-bound chains, not throughput:
    a GPU advances every resident hidden-layer model one step per chain_latency_us(count) and its single-Dense pack
    concurrently on the SMs the clusters leave free (vae.train_many: one stream per kernel class).  So the
    hidden-layer runs are dealt out evenly first (longest first, to the rank with the smallest resulting
    steps x latency), then the single-Dense runs by LPT over steps, preferring the ranks with the fewest clusters --
    instead of one LPT over flops, which parks all 39 small runs beside a sphere chain on one GPU (round 1: 27.7 s
    on the slowest of 8 ranks against a 22 s chain)."""
    hidden = [i for i, a in enumerate(all_args) if _spec_of(a).layer_sizes or _spec_of(a).encoder_layer_sizes]
    small = [i for i in range(len(all_args)) if i not in set(hidden)]
    parts = [[] for _ in range(n_parts)]
    n_hidden = [0] * n_parts
    steps_hidden = [0] * n_parts
    for i in sorted(hidden, key=lambda i: -all_args[i].num_batches):
        def wall(q):
            return max(steps_hidden[q], all_args[i].num_batches) * chain_latency_us(n_hidden[q] + 1)
        p = min(range(n_parts), key=lambda q: (wall(q), n_hidden[q], q))
        parts[p].append(i)
        n_hidden[p] += 1
        steps_hidden[p] = max(steps_hidden[p], all_args[i].num_batches)
    load = [0.0] * n_parts
    for i in sorted(small, key=lambda i: -run_cost(all_args[i])):
        p = min(range(n_parts), key=lambda q: (load[q], n_hidden[q], q))
        parts[p].append(i)
        load[p] += run_cost(all_args[i])
    return parts


def estimated_wall_seconds(all_args, part) -> float:
    """The model latency_partition minimises, for one rank's runs (reported next to measured times)."""
    hidden = [i for i in part if _spec_of(all_args[i]).layer_sizes or _spec_of(all_args[i]).encoder_layer_sizes]
    small = [i for i in part if i not in set(hidden)]
    t2 = max((all_args[i].num_batches for i in hidden), default=0) * chain_latency_us(len(hidden)) * 1e-6
    t1 = max((all_args[i].num_batches for i in small), default=0) * _SMALL_STEP_US * 1e-6
    return max(t1, t2)


def run_sweep(all_args, rank=0, world=1, save=True, verbose=False, timings=None):
    import time

    import torch

    from .vae import train_many
    t0 = time.perf_counter()
    mine = latency_partition(all_args, world)[rank]
    models = []
    for i in mine:
        a = all_args[i]
        out_dir = runmod.make_output_dir(a.name, a.overwrite, a) if save else None
        ds = runmod.get_dataset(a.dataset, a.dataset_seed, a.padding_dim, a.batch_size, a)
        models.append(runmod.get_model(a, ds, out_dir, verbose=verbose))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    train_many(models, save=save, timings=timings)
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
