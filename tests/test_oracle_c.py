"""The plain-C oracle (oracle/c, the timed CPU baseline) against the NumPy oracle and the
published known answers: RNG / generators bit-exact, step within float32 tolerance."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import datasets as ods
from oracle import jax_random as jr
from oracle import trainer as otr
from oracle import vae_step as vs


def test_c_threefry_kat_and_documented_values():
    assert co.split(jr.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert np.float32(co.normal(jr.PRNGKey(0), ())) == np.float32(-0.20584226)
    assert np.float32(co.normal(jr.PRNGKey(42), ())) == np.float32(-0.18471177)
    import ctypes as C
    x0, x1 = C.c_uint32(0x243f6a88), C.c_uint32(0x85a308d3)
    co.load().or_threefry2x32(C.c_uint32(0x13198a2e), C.c_uint32(0x03707344), C.byref(x0), C.byref(x1))
    assert (x0.value, x1.value) == (0xc4923a9c, 0x483df7a0)


@pytest.mark.parametrize("n", [1, 2, 5, 1300, 5901, 200001])
def test_c_bits_and_normal_bit_exact(n):
    key = jr.PRNGKey(77 + n)
    np.testing.assert_array_equal(co.random_bits(key, n), jr.random_bits(key, n))
    np.testing.assert_array_equal(co.normal(key, (n,)).view(np.uint32), jr.normal(key, (n,)).view(np.uint32))


@pytest.mark.parametrize("name,dd,pd", [("sphere", 3, 3), ("linear_gaussian", 6, 14), ("sigmoid", 7, 20)])
def test_c_dataset_bit_exact(name, dd, pd):
    ds = ods.make_dataset(name, 48, dd, pd)
    key = ds.key.copy()
    ref = ds.get_batch(100)
    bkey, _ = jr.dataset_chain_next(key)
    got = co.dataset_batch(co.make_data(name, dd, pd), bkey, ds.A, 100)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


CASES = [vs.VaeDesc(100, 12, 20, epsilon0=-1.0, lr=1e-3),
         vs.VaeDesc(100, 7, 6, sig_decoder=True, epsilon0=-3.0),
         vs.VaeDesc(100, 6, 6, [200, 200, 200], [200, 200, 200], epsilon0=-3.0),
         vs.VaeDesc(37, 5, 3, [17], [9, 33], tunable_decoder_var=False, epsilon0=-0.5)]


@pytest.mark.parametrize("d", CASES, ids=lambda d: f"D{d.data_dim}_L{d.latent_dim}_h{len(d.enc_hidden)}")
def test_c_loss_and_grad(d):
    rng = np.random.default_rng(0)
    flat = (vs.init_params(d, 1) + 0.02 * rng.standard_normal(d.n_params)).astype(np.float32)
    x = rng.standard_normal((d.batch, d.data_dim)).astype(np.float32)
    z = rng.standard_normal((d.batch, d.latent_dim + d.data_dim)).astype(np.float32)
    L = d.latent_dim
    l64, g64 = vs.loss_and_grad(d, flat.astype(np.float64), x, z[:, :L], z[:, L:], dtype=np.float64)
    loss, grad = co.loss_and_grad(co.make_desc(d), flat, x, z)
    assert abs(loss - l64) <= 1e-5 * abs(l64)
    assert np.max(np.abs(grad - g64)) <= 1e-5 * np.max(np.abs(g64))


@pytest.mark.parametrize("name,dd,pd,L,hidden", [("sigmoid", 3, 3, 6, []), ("linear_gaussian", 3, 9, 20, []),
                                                 ("sphere", 3, 3, 6, [200, 200, 200])])
def test_c_train_loop_matches_numpy_oracle(name, dd, pd, L, hidden):
    ds = ods.make_dataset(name, 69, dd, pd)
    d = vs.VaeDesc(100, ds.ndim, L, hidden, hidden, sig_decoder=(name == "sigmoid"), epsilon0=-3.0, lr=1e-4)
    flat = vs.init_params(d, 0)
    tr = otr.OracleTrainer(d, ds, flat, dtype=np.float64)
    tr.begin()
    n = 12
    job = co.Job(co.make_desc(d), co.make_data(name, dd, pd), flat, tr.key, ds.key, ds.A, n)
    co.train_loop([job], n)
    tr.train(10 ** 9, 0, n, events=False)
    np.testing.assert_array_equal(job.model_key, tr.key)
    np.testing.assert_array_equal(job.dataset_key, ds.key)
    ref = np.array(tr.train_losses)
    assert np.max(np.abs(job.loss[:n] - ref) / np.abs(ref)) <= 1e-5
    diff = np.abs(job.params - tr.params)
    assert np.quantile(diff, 0.999) <= n * 1e-3 * d.lr + 1e-7 and diff.max() <= 0.5 * n * d.lr
    assert int(job.step[0]) == n


def test_c_multithreaded_equals_single():
    ds = ods.make_dataset("sigmoid", 69, 3, 3)
    d = vs.VaeDesc(100, 7, 6, sig_decoder=True, epsilon0=-3.0)
    flat = vs.init_params(d, 0)
    mk = jr.PRNGKey(1)
    mk_jobs = lambda: [co.Job(co.make_desc(d), co.make_data("sigmoid", 3, 3), flat, mk, jr.PRNGKey(s), ds.A, 20)
                       for s in range(6)]
    a, b = mk_jobs(), mk_jobs()
    co.train_loop(a, 20, n_threads=1)
    co.train_loop(b, 20, n_threads=4)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x.params, y.params)
        np.testing.assert_array_equal(x.loss, y.loss)
