"""refshim flax.optim.Adam (legacy; vae.py:113, networks.py:100).  Update rule restated:
t = step + 1; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps);
defaults b1 = 0.9, b2 = 0.999, eps = 1e-8, weight_decay = 0.  float64 throughout."""
import torch


def _map(fn, *trees):
    t0 = trees[0]
    return {k: _map(fn, *[t[k] for t in trees]) if isinstance(v, dict) else fn(*[t[k] for t in trees])
            for k, v in t0.items()}


class Optimizer:
    def __init__(self, optimizer_def, target, step, grad_ema, grad_sq_ema):
        self.optimizer_def, self.target, self.step = optimizer_def, target, step
        self.grad_ema, self.grad_sq_ema = grad_ema, grad_sq_ema

    def apply_gradient(self, grads, **hyper_param_overrides):
        from .nn import Model
        h = self.optimizer_def
        g = grads.params if isinstance(grads, Model) else grads
        t = self.step + 1.0
        m = _map(lambda m_, g_: h.beta1 * m_ + (1.0 - h.beta1) * g_, self.grad_ema, g)
        v = _map(lambda v_, g_: h.beta2 * v_ + (1.0 - h.beta2) * g_ * g_, self.grad_sq_ema, g)
        bc1, bc2 = 1.0 - h.beta1 ** t, 1.0 - h.beta2 ** t

        def upd(p, m_, v_):
            new = p - h.learning_rate * (m_ / bc1) / (torch.sqrt(v_ / bc2) + h.eps)
            return new - h.learning_rate * h.weight_decay * p
        params = _map(upd, self.target.params, m, v)
        return Optimizer(h, Model(self.target.module, params), self.step + 1, m, v)


class Adam:
    def __init__(self, learning_rate=None, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        self.learning_rate, self.beta1, self.beta2, self.eps, self.weight_decay = learning_rate, beta1, beta2, eps, weight_decay

    def create(self, target):
        zeros = _map(lambda p: torch.zeros_like(p), target.params)
        zeros2 = _map(lambda p: torch.zeros_like(p), target.params)
        return Optimizer(self, target, 0, zeros, zeros2)
