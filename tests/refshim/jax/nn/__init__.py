"""refshim jax.nn."""
import torch

from . import initializers  # noqa: F401


def sigmoid(x):
    return 1.0 / (1.0 + torch.exp(-x))
