"""The 54 runs of the reference's three sweep scripts, regenerated from their parameter grids
(seed_linpadding_expts.sh:1-21, sigmoid_vae_padding_expts.sh:1-20,
sphere_vae_padding_expts.sh:1-17) as run.py argument lists, so the
full sweep can be launched where the reference tree is not available (the GPU box).

Every run: batch 100, ``-tdv -ow``.  Seeds vary only the DATASET key (-ds); the model key is
always PRNGKey(0) (model.py:29)."""
from __future__ import annotations

from typing import List

# (dd, pd, latent_dim)
LINEAR_GRID = [(3, 9, 20), (3, 17, 20), (6, 6, 20), (6, 14, 20), (9, 3, 20), (9, 11, 10), (12, 8, 10)]
SIGMOID_GRID = [(3, 3, 6), (3, 13, 8), (5, 16, 16), (5, 5, 10), (7, 7, 13), (7, 20, 24)]
SPHERE_GRID = [(3, 3, 6), (3, 13, 8), (5, 16, 16), (5, 5, 10), (7, 7, 13)]
LINEAR_SEEDS = (2, 3, 4)
OTHER_SEEDS = (69, 24, 48)


def linear_runs() -> List[List[str]]:
    out = []
    for ds in LINEAR_SEEDS:
        for dd, pd, lat in LINEAR_GRID:
            out.append([f"vae{dd}linear_gaussian_{dd + pd}dim{ds}", "--dataset", "linear_gaussian",
                        "--encoder_layer_sizes", "", "--layer_sizes", "", "-ow", "--latent_dim", str(lat),
                        "--padding_dim", str(pd), "-dd", str(dd), "--num_batches", "100000", "--epsilon", "-1", "-tdv",
                        "-ds", str(ds), "-lr", "1e-3"])
    return out


def _padding_runs(dataset: str, grid, hidden: str) -> List[List[str]]:
    out = []
    for ds in OTHER_SEEDS:
        for dd, pd, lat in grid:
            # names and flags exactly as the scripts spell them (..._seed24, --dataset_seed 24)
            name = f"{dataset}_dd{dd}_pd{pd}_ld_{lat}_eps-3" + ("" if ds == 69 else f"_seed{ds}")
            args = [name, "--dataset", dataset, "--encoder_layer_sizes", hidden, "--layer_sizes", hidden, "-ow",
                    "--latent_dim", str(lat), "--padding_dim", str(pd), "-dd", str(dd), "--num_batches", "150000",
                    "--epsilon", "-3", "-tdv"]
            if ds != 69:
                args += ["--dataset_seed", str(ds)]
            out.append(args)
    return out


def sigmoid_runs() -> List[List[str]]:
    return _padding_runs("sigmoid", SIGMOID_GRID, "")


def sphere_runs() -> List[List[str]]:
    return _padding_runs("sphere", SPHERE_GRID, "200|200|200")


def all_runs() -> List[List[str]]:
    return linear_runs() + sigmoid_runs() + sphere_runs()


BUILTIN = {"linear": linear_runs, "sigmoid": sigmoid_runs, "sphere": sphere_runs, "all": all_runs}
