"""Checks oracle/vae_step.py (hand-derived backward + legacy flax Adam) against torch
autograd in float64, and the generators / event schedule of oracle/datasets.py and
oracle/trainer.py against their definitions in the reference (datasets.py:75-84,
183-195, 240-249; model.py:207-222)."""
import numpy as np
import pytest
import torch

from oracle import datasets as ods
from oracle import jax_random as jr
from oracle import trainer as otr
from oracle import vae_step as vs

CASES = {
    "linear": vs.VaeDesc(batch=100, data_dim=12, latent_dim=20, epsilon0=-1.0, lr=1e-3),
    "sigmoid": vs.VaeDesc(batch=100, data_dim=7, latent_dim=6, sig_decoder=True, epsilon0=-3.0),
    "sphere": vs.VaeDesc(batch=64, data_dim=6, latent_dim=6, enc_hidden=[40, 30, 20], dec_hidden=[20, 30, 40],
                         epsilon0=-3.0),
    "fixed_var": vs.VaeDesc(batch=32, data_dim=5, latent_dim=3, enc_hidden=[16], dec_hidden=[16],
                            tunable_decoder_var=False, epsilon0=-0.5),
}


def torch_loss(desc, flat, x, z1, z2):
    """networks.py:61-98 written directly in torch (float64) for autograd."""
    p = {name: flat[off:off + int(np.prod(shape))].reshape(shape) for name, shape, off in desc.segments()}

    def mlp(net, sizes, h):
        n = len(sizes) - 1
        for i in range(n):
            h = h @ p[f"{net}/FC{i}/kernel"] + p[f"{net}/FC{i}/bias"]
            if i + 1 < n:
                h = torch.relu(h)
        return h

    mu = mlp("Encoder", desc.enc_sizes, x)
    lv = p["epsilon_p"]
    e = p["epsilon"] * desc.epsilon0 if desc.tunable_decoder_var else torch.full((1,), desc.epsilon0, dtype=x.dtype)
    zs = mu + torch.exp(lv / 2) * z1
    xhat = mlp("Decoder", desc.dec_sizes, zs)
    if desc.sig_decoder:
        xhat = torch.sigmoid(mlp("SigDecoder", desc.dec_sizes, zs)) + xhat
    xhat = xhat + z2 * torch.exp(e / 2)
    dkl = -0.5 * torch.sum(1 + lv - torch.exp(lv) - mu ** 2, dim=-1)
    mse = (0.5 * (xhat - x) ** 2 / torch.exp(e) + 0.5 * (np.log(2 * np.pi) + e)).sum(dim=-1)
    return (dkl + mse).mean()


@pytest.mark.parametrize("name", list(CASES))
def test_backward_matches_autograd_f64(name):
    desc = CASES[name]
    rng = np.random.default_rng(1)
    flat = vs.init_params(desc, seed=3, dtype=np.float64)
    flat += 0.05 * rng.standard_normal(flat.size)          # move biases / eps off their init
    x = rng.standard_normal((desc.batch, desc.data_dim))
    z1 = rng.standard_normal((desc.batch, desc.latent_dim))
    z2 = rng.standard_normal((desc.batch, desc.data_dim))
    loss, grad = vs.loss_and_grad(desc, flat, x, z1, z2, dtype=np.float64)
    t = torch.tensor(flat, dtype=torch.float64, requires_grad=True)
    tl = torch_loss(desc, t, torch.tensor(x), torch.tensor(z1), torch.tensor(z2))
    tl.backward()
    assert abs(float(tl) - float(loss)) <= 1e-12 * abs(float(tl))
    tg = t.grad.numpy()
    assert np.max(np.abs(grad - tg)) <= 1e-12 * max(1.0, np.max(np.abs(tg)))


def test_adam_matches_reference_formula():
    """Legacy flax.optim.Adam: eps added after sqrt of the bias-corrected v."""
    desc = CASES["linear"]
    rng = np.random.default_rng(0)
    p = rng.standard_normal(desc.n_params)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    pt = torch.tensor(p.copy(), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([pt], lr=desc.lr, betas=(0.9, 0.999), eps=1e-8)
    step = 0
    for _ in range(5):
        g = rng.standard_normal(desc.n_params)
        p, m, v, step = vs.adam_update(desc, p, m, v, step, g, dtype=np.float64)
        pt.grad = torch.tensor(g)
        opt.step()          # torch: p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)  -- same algebra
    np.testing.assert_allclose(p, pt.detach().numpy(), rtol=1e-10, atol=1e-12)
    assert step == 5


def test_f32_step_close_to_f64_shadow():
    desc = CASES["sigmoid"]
    rng = np.random.default_rng(5)
    flat = vs.init_params(desc, seed=1)
    x = rng.standard_normal((desc.batch, desc.data_dim)).astype(np.float32)
    z = rng.standard_normal((desc.batch, desc.latent_dim + desc.data_dim)).astype(np.float32)
    L = desc.latent_dim
    l32, g32 = vs.loss_and_grad(desc, flat, x, z[:, :L], z[:, L:])
    l64, g64 = vs.loss_and_grad(desc, flat.astype(np.float64), x, z[:, :L], z[:, L:], dtype=np.float64)
    assert abs(l32 - l64) <= 2e-6 * abs(l64)
    assert np.max(np.abs(g32 - g64)) <= 2e-5 * np.max(np.abs(g64))


def test_param_counts_match_survey_appendix_c():
    assert vs.VaeDesc(100, 12, 20).n_params == 533
    assert vs.VaeDesc(100, 20, 20).n_params == 861
    assert vs.VaeDesc(100, 20, 10).n_params == 441
    assert vs.VaeDesc(100, 7, 6, sig_decoder=True).n_params == 153
    assert vs.VaeDesc(100, 28, 24, sig_decoder=True).n_params == 2121
    h = [200, 200, 200]
    assert vs.VaeDesc(100, 6, 6, h, h).n_params == 166019
    assert vs.VaeDesc(100, 21, 16, h, h).n_params == 176054
    assert vs.VaeDesc(100, 12, 20).flops_per_step() == 240000
    assert vs.VaeDesc(100, 7, 6, sig_decoder=True).flops_per_step() == 67200
    assert vs.VaeDesc(100, 6, 6, h, h).flops_per_step() == 98640000


# ------------------------------------------------------------------ generators

def test_sphere_batch_definition():
    ds = ods.SphereDataset(69, 3, 3)
    k0 = ds.key.copy()
    batch = ds.get_batch(100)
    key, carried = jr.dataset_chain_next(k0)
    assert ds.key.tolist() == carried.tolist()
    g = jr.normal(key, (100, 3)).astype(np.float64)
    ref = g / np.linalg.norm(g, axis=1, keepdims=True)
    assert batch.shape == (100, 6) and batch.dtype == np.float32
    np.testing.assert_allclose(batch[:, :3], ref, rtol=3e-7)
    assert (batch[:, 3:] == 0).all()
    sc = ds.score_batch(batch)
    assert sc["Sphere Error"] < 1e-12 and sc["Padding Error"] == 0


def test_linear_batch_definition():
    ds = ods.LinearGaussianDataset(2, 3, 3, 9)
    assert ds.A.shape == (3, 3)
    ret, _ = jr.dataset_chain_next(jr.PRNGKey(2))
    np.testing.assert_array_equal(ds.A, jr.normal(ret, (3, 3)))         # datasets.py:172-173
    k0 = ds.key.copy()
    batch = ds.get_batch(100)
    key, _ = jr.dataset_chain_next(k0)
    X = jr.normal(key, (100, 3)).astype(np.float64)
    np.testing.assert_allclose(batch[:, :3], (ds.A.astype(np.float64) @ X.T).T, rtol=1e-5, atol=1e-6)
    assert batch.shape == (100, 12) and (batch[:, 3:] == 0).all()


def test_sigmoid_batch_definition():
    ds = ods.SigmoidDataset(69, 3, 3)
    k0 = ds.key.copy()
    batch = ds.get_batch(100)
    key, _ = jr.dataset_chain_next(k0)
    z = jr.normal(key, (100, 3))
    np.testing.assert_array_equal(batch[:, :3], z)
    ref = 1 / (1 + np.exp(-(z.astype(np.float64) @ ds.A.astype(np.float64))))
    np.testing.assert_allclose(batch[:, 3], ref[:, 0], rtol=3e-7)
    assert batch.shape == (100, 7) and (batch[:, 4:] == 0).all()


def test_det_expf_accuracy():
    x = np.linspace(-30, 30, 100001).astype(np.float32)
    ref = np.exp(x.astype(np.float64))
    got = ods.det_expf(x).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) <= 2.5e-7


# ------------------------------------------------------------------ schedule

def test_event_schedule_key_chain():
    """model.py:207-222 with n_print / n_plot shrunk: every model-side split and every
    dataset-side get_key is accounted for."""
    desc = CASES["sigmoid"]
    ds = ods.SigmoidDataset(69, 3, 3)
    tr = otr.OracleTrainer(desc, ds, vs.init_params(desc, 0), n_print=4, n_plot=8, eval_batch=10)
    n = 10
    tr.train(n)
    # replay the chains by hand
    key = jr.PRNGKey(0)
    _, key = jr.model_chain_next(key)        # vae.py:56
    _, key = jr.model_chain_next(key)        # model.py:208
    dkey = jr.PRNGKey(69)
    _, dkey = jr.dataset_chain_next(dkey)    # ctor A
    _, dkey = jr.dataset_chain_next(dkey)    # model.py:209
    lat = []
    for i in range(n):
        if i % 4 == 0:
            _, key = jr.model_chain_next(key)
            _, dkey = jr.dataset_chain_next(dkey)
        if i % 8 == 0 or i == n - 1:
            _, key = jr.model_chain_next(key)
            _, dkey = jr.dataset_chain_next(dkey)       # sigmoid plot_batch draws (datasets.py:265)
        _, dkey = jr.dataset_chain_next(dkey)
        k, key = jr.model_chain_next(key)
        lat.append(k)
    assert tr.key.tolist() == key.tolist()
    assert ds.key.tolist() == dkey.tolist()
    assert [k.tolist() for k in tr.latent_keys] == [k.tolist() for k in lat]
    assert len(tr.vae_losses) == n + 3 and len(tr.train_losses) == n     # vae.py:130,136
    assert len(tr.var_enc) == 3 and tr.var_enc[0].shape == (6,) and tr.var_dec[0].shape == (1,)


def test_first_latent_key_of_a_real_run_matches_survey():
    desc = CASES["sigmoid"]
    ds = ods.SigmoidDataset(69, 3, 3)
    tr = otr.OracleTrainer(desc, ds, vs.init_params(desc, 0))
    tr.train(150000, start=0, stop=2)
    assert tr.latent_keys[0].tolist() == [145021652, 2649282573]
    assert tr.latent_keys[1].tolist() == [1409544366, 3423741852]
    assert np.isfinite(tr.train_losses).all()
