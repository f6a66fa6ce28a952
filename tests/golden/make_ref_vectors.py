"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE, unmodified.

    python tests/golden/make_ref_vectors.py            # needs /root/reference (this container only)

`/root/reference/{run,vae,model,networks,datasets,utils}.py` are imported via sys.path on top of
tests/refshim (a minimal jax / legacy-flax API surface: float64 torch + autograd, jax.random =
oracle/jax_random.py; see tests/refshim/README.md) and driven through `run.py`'s own
`parse_arguments` / `main` with the flags of the sweep scripts.  Nothing of the reference is copied
or patched; the only environment fixes are `numpy.product = numpy.prod` (removed in NumPy 2) and a
no-op matplotlib.

Recorded per run: the flags, the initial parameters (the shim's init -- an INPUT, see refshim),
`vae_losses` (train and eval losses interleaved exactly as vae.py:130,136 append them), every stat
list of model.py:195-205, `var_enc` / `var_dec`, both key chains at the end, final parameters and
Adam moments, and the generator matrix A.  Runs use the reference's real cadence (n_print = 5000,
n_plot = 50000) for > 2 event cycles.
"""
import argparse
import importlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

RUNS = {
    # name: argv of `python run.py ...` (the first line of each sweep script, shorter num_batches)
    "ref_linear_dd3_pd9_ld20": ["vae3linear_gaussian_12dim2", "--dataset", "linear_gaussian", "--encoder_layer_sizes", "",
                                "--layer_sizes", "", "-ow", "--latent_dim", "20", "--padding_dim", "9", "-dd", "3",
                                "--num_batches", "10050", "--epsilon", "-1", "-tdv", "-ds", "2", "-lr", "1e-3"],
    "ref_sigmoid_dd3_pd3_ld6": ["sigmoid_dd3_pd3_ld_6_eps-3", "--dataset", "sigmoid", "--encoder_layer_sizes", "",
                                "--layer_sizes", "", "-ow", "--latent_dim", "6", "--padding_dim", "3", "-dd", "3",
                                "--num_batches", "10050", "--epsilon", "-3", "-tdv"],
    "ref_sigmoid_dd5_pd5_ld10_seed24": ["sigmoid_dd5_pd5_ld_10_eps-3_seed24", "--dataset", "sigmoid",
                                        "--encoder_layer_sizes", "", "--layer_sizes", "", "-ow", "--latent_dim", "10",
                                        "--padding_dim", "5", "-dd", "5", "--num_batches", "5003", "--epsilon", "-3",
                                        "-tdv", "--dataset_seed", "24"],
    "ref_sphere_dd3_pd3_ld6": ["sphere_dd3_pd3_ld_6_eps-3", "--dataset", "sphere", "--encoder_layer_sizes", "200|200|200",
                               "--layer_sizes", "200|200|200", "-ow", "--latent_dim", "6", "--padding_dim", "3", "-dd", "3",
                               "--num_batches", "5003", "--epsilon", "-3", "-tdv"],
    # fixed decoder variance (no -tdv) and a noisy linear dataset (-dn): flag coverage beyond the scripts
    "ref_sphere_small_fixedvar": ["sphere_small", "--dataset", "sphere", "--encoder_layer_sizes", "32|16",
                                  "--layer_sizes", "24", "-ow", "--latent_dim", "4", "--padding_dim", "2", "-dd", "4",
                                  "--num_batches", "120", "--epsilon", "-2", "--batch_size", "37"],
    "ref_linear_noise_dn": ["lin_noise", "--dataset", "linear_gaussian", "--encoder_layer_sizes", "", "--layer_sizes", "",
                            "-ow", "--latent_dim", "5", "--padding_dim", "2", "-dd", "4", "--num_batches", "60",
                            "--epsilon", "-1", "-tdv", "-ds", "3", "-lr", "1e-3", "-dn", "0.25"],
    # warm start (-ws, vae.py:61-103): the hand-built starting point + perturbations drawn with the carried model key
    "ref_sigmoid_warm_start": ["sig_ws", "--dataset", "sigmoid", "--encoder_layer_sizes", "", "--layer_sizes", "", "-ow",
                               "--latent_dim", "6", "--padding_dim", "2", "-dd", "3", "--num_batches", "40", "--epsilon", "-3",
                               "-tdv", "-ws"],
    "ref_linear_warm_start": ["lin_ws", "--dataset", "linear_gaussian", "--encoder_layer_sizes", "", "--layer_sizes", "",
                              "-ow", "--latent_dim", "8", "--padding_dim", "3", "-dd", "3", "--num_batches", "40",
                              "--epsilon", "-1", "-tdv", "-ds", "2", "-lr", "1e-3", "-ws", "-off", "1"],
}


def import_reference():
    """sys.path: refshim (jax, flax, matplotlib), the repo root (oracle), then the reference tree."""
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} is not available: the fixtures can only be regenerated where the reference tree is")
    for p in (os.path.join(ROOT, "tests", "refshim"), ROOT, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference's module names are generic (`datasets`, `model`, `run`, `utils`, `vae`): make sure they are its own
    for name in ("run", "vae", "model", "networks", "datasets", "utils"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF):
            del sys.modules[name]
    if not hasattr(np, "product"):
        np.product = np.prod                      # datasets.py:28, vae.py:51 (NumPy < 2 spelling)
    sys.path.insert(0, REF)
    run = importlib.import_module("run")
    assert run.__file__.startswith(REF)
    return run


def flatten_params(tree, names):
    """Reference parameter tree -> the flat layout of include/vae_b200.h (Encoder, Decoder, [SigDecoder],
    epsilon_p, [epsilon]; per layer kernel (in,out) row-major then bias)."""
    parts = []
    for net in ("Encoder", "Decoder", "SigDecoder"):
        if net in tree:
            for i in range(len(tree[net])):
                parts.append(tree[net][f"FC{i}"]["kernel"].detach().numpy().ravel())
                parts.append(tree[net][f"FC{i}"]["bias"].detach().numpy().ravel())
    parts.append(tree["epsilon_p"].detach().numpy().ravel())
    if "epsilon" in tree:
        parts.append(tree["epsilon"].detach().numpy().ravel())
    return np.concatenate(parts).astype(np.float64)


def run_reference(run, argv):
    """`python run.py argv...` exactly as run.py:363-369 does, in a scratch cwd (it writes data/<name>/)."""
    old_argv, old_cwd = sys.argv, os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, "data"))
        os.chdir(tmp)
        sys.argv = ["run.py"] + list(argv)
        try:
            args = run.parse_arguments()
            args.tqdm = False
            # run.py:350-360 (main), kept step by step so the objects can be inspected afterwards
            output_dir = run.make_output_dir(args.name, args.overwrite, args)
            dataset = run.get_dataset(args.dataset, args.dataset_seed, args.padding_dim, args.batch_size, args)
            model = run.get_model(args, dataset, output_dir)
            init = flatten_params(model.model.params, None)
            A = None if getattr(dataset, "A", None) is None else dataset.A.numpy().copy()
            model.train()
            model.plot()
            model.save(final=True)
            saved = dict(np.load(os.path.join(output_dir, "losses.npz"), allow_pickle=True))
            files = sorted(os.listdir(output_dir))
        finally:
            sys.argv = old_argv
            os.chdir(old_cwd)
    opt = model.optimizer
    out = {
        "argv": np.array(argv),
        "init_params": init,
        "A": np.zeros(0) if A is None else A,
        "vae_losses": np.array([float(v) for v in model.vae_losses]),
        "var_enc": np.array([v.detach().numpy() for v in model.var_enc]),
        "var_dec": np.array([np.atleast_1d(v.detach().numpy() if hasattr(v, "detach") else v) for v in model.var_dec]),
        "final_params": flatten_params(opt.target.params, None),
        "final_m": flatten_params(opt.grad_ema, None),
        "final_v": flatten_params(opt.grad_sq_ema, None),
        "final_step": np.array(opt.step),
        "model_key": np.asarray(model.key, np.uint32),
        "dataset_key": np.asarray(dataset.key, np.uint32),
        "current_epsilon": np.atleast_1d(np.asarray(model.current_epsilon, np.float64)),
        "saved_keys": np.array(sorted(saved)),
        "saved_vae_loss_len": np.array(len(saved["VAE Loss"])),
        "output_files": np.array(files),
    }
    for stat, vals in model.stats.items():
        out["stat/" + stat] = np.array([float(v) for v in vals])
    if init.size > 20000:
        # keep the fixture small: initial parameters are float32-exact (stored as float32); the final state of a
        # 166k-parameter model is kept on a fixed stride subsample (+ its index) and as whole-vector norms
        idx = np.arange(0, init.size, 41)
        out["init_params"] = init.astype(np.float32)
        out["final_index"] = idx
        for k in ("final_params", "final_m", "final_v"):
            out[k + "_norm"] = np.array(np.linalg.norm(out[k]))
            out[k] = out[k][idx]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    run = import_reference()
    for name, argv in RUNS.items():
        if a.only and a.only != name:
            continue
        out = run_reference(run, argv)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "steps", int(out["final_step"]), "losses", out["vae_losses"].shape, "first", out["vae_losses"][:3],
              "last", out["vae_losses"][-1], flush=True)


if __name__ == "__main__":
    main()
