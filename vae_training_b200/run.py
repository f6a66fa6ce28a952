"""Flag-compatible front end: the argparse surface of the reference's run.py:8-43, its
dataset / model factories (run.py:46-54, 250-268) and main (run.py:350-360).

    python -m vae_training_b200.run NAME --dataset sigmoid -dd 3 --padding_dim 3 --latent_dim 6 \
        --encoder_layer_sizes "" --layer_sizes "" --num_batches 150000 --epsilon -3 -tdv -ow
"""
from __future__ import annotations

import argparse
import json
import os

DATA_DIR = "data"      # utils.py:11


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    parser.add_argument('name', help="The name of the experiment and output directory.")
    parser.add_argument('--num_batches', dest='num_batches', type=int, default=15000,
                        help="Number of batches to train on.")
    parser.add_argument('--num_epochs', dest='num_epochs', type=int, default=10000)
    parser.add_argument('--batch_size', dest='batch_size', type=int, default=100)
    parser.add_argument('-lr', '--learning_rate', dest='learning_rate', type=float, default=0.0001)
    parser.add_argument('--padding_dim', type=int, dest='padding_dim', default=0)
    parser.add_argument('-ow', dest='overwrite', action='store_true')
    parser.add_argument('--dataset', dest='dataset', default='4gaussian',
                        choices=["sphere", "linear_gaussian", "sigmoid"])
    parser.add_argument('--layer_sizes', dest='layer_sizes', default='512|512')
    parser.add_argument('--encoder_layer_sizes', dest='encoder_layer_sizes', default='512|512')
    parser.add_argument('--latent_dim', dest='latent_dimension', type=int, default=100)
    parser.add_argument('-nojit', dest='nojit', action='store_true')
    parser.add_argument('--padding_type', dest='padding_type', default="none", choices=["zero", "gaussian", "none"])
    parser.add_argument('-ds', '--dataset_seed', dest='dataset_seed', type=int, default=69)
    parser.add_argument('--state_dict', dest='state_dict', default=None)
    parser.add_argument('--data_fn', dest="data_fn", default=None)
    parser.add_argument('-ws', '--warm_start', action='store_true')
    parser.add_argument('-ii', '--initialize_inverse', action='store_true')
    parser.add_argument('-ufc', '--use_fred_covariance', action='store_true')
    parser.add_argument('-e', '--epsilon', type=float, default=0.)
    parser.add_argument('-tdv', dest='tunable_decoder_var', action='store_true')
    parser.add_argument('-dn', '--dataset_noise', type=float, default=0.)
    parser.add_argument('-dd', '--dataset_dimension', type=int, default=3)
    parser.add_argument('-wsl', '--warm_start_linear', action='store_true')
    parser.add_argument('-did', '--dataset_intrinsic_dimension', type=int, default=3)
    parser.add_argument('-off', '--latent_off_dimension', type=int, default=1)
    return parser


def parse_arguments(argv=None):
    args = build_parser().parse_args(argv)
    args.model = "VAE"                      # run.py:40-42
    args.latent_distribution = 'gaussian'
    args.tqdm = True
    return args


def make_output_dir(name, overwrite, args, data_dir=None):      # utils.py:46-60
    dirname = os.path.join(data_dir or DATA_DIR, name)
    if os.path.exists(dirname):
        if overwrite:
            for fn in os.listdir(dirname):
                os.remove(os.path.join(dirname, fn))
        else:
            raise ValueError(f"{dirname} already exists! Use a different name")
    else:
        os.makedirs(dirname)
    with open(os.path.join(dirname, 'args.json'), 'w') as f:
        json.dump(vars(args), f)
    return dirname


def get_dataset(name, seed, padding_dimension, batch_size, args):   # run.py:46-54
    from .datasets import LinearGaussianDataset, SigmoidDataset, SphereDataset
    if name == "sphere":
        return SphereDataset(seed, dimension=args.dataset_dimension, padding_dimension=args.padding_dim)
    elif name == "linear_gaussian":
        return LinearGaussianDataset(seed, dimension=args.dataset_dimension,
                                     intrinsic_dimension=args.dataset_intrinsic_dimension,
                                     padding_dimension=args.padding_dim, var_added=args.dataset_noise)
    elif name == "sigmoid":
        return SigmoidDataset(seed, dimension=args.dataset_dimension, padding_dimension=args.padding_dim)
    raise ValueError(f"unknown dataset {name!r}")


def get_model(args, dataset, output_dir, verbose=True):             # run.py:250-268
    from .vae import VAEModel
    return VAEModel(dirname=output_dir, batch_size=args.batch_size, learning_rate=args.learning_rate,
                    dataset=dataset, num_batches=args.num_batches, num_epochs=args.num_epochs,
                    layer_sizes=args.layer_sizes, encoder_layer_sizes=args.encoder_layer_sizes,
                    state_dict=args.state_dict, data_fn=args.data_fn, epsilon=args.epsilon, tqdm=args.tqdm,
                    latent_dimension=args.latent_dimension, tunable_decoder_var=args.tunable_decoder_var,
                    warm_start=args.warm_start, dataset_name=args.dataset,
                    latent_off_dimension=args.latent_off_dimension, verbose=verbose)


def main(args):                                                     # run.py:350-360
    output_dir = make_output_dir(args.name, args.overwrite, args)
    dataset = get_dataset(args.dataset, args.dataset_seed, args.padding_dim, args.batch_size, args)
    model = get_model(args, dataset, output_dir)
    model.train()
    model.plot()
    model.save(final=True)
    return 0


if __name__ == '__main__':
    main(parse_arguments())
