"""refshim jax.scipy.stats: imported (networks.py:2, model.py:4, datasets.py:5), never called on the VAE path."""


class _Unavailable:
    def __getattr__(self, name):
        raise NotImplementedError("refshim: jax.scipy.stats is not on the VAE path")


norm = _Unavailable()
logistic = _Unavailable()
multivariate_normal = _Unavailable()
