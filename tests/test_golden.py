"""Golden vectors (tests/golden/): the oracle -- and, under -m gpu, the CUDA path through the
C ABI -- against committed known answers.

* jax_random_kat.json  published Random123 / jax.random values + the derived key-chain values of
                       SURVEY.md A.5 (independent of oracle/).
* step_vectors.npz     regression vectors for the step (loss, gradient, Adam update) produced by
                       tests/golden/make_step_vectors.py from the float64 oracle, which is itself
                       pinned by torch autograd (tests/test_oracle_step.py).  The reference cannot
                       be executed here, so these are NOT reference outputs (see the file header).
"""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import jax_random as jr
from oracle import vae_step as vs

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "jax_random_kat.json")))
STEP = np.load(os.path.join(HERE, "golden", "step_vectors.npz"))
U32 = np.uint32


def step_cases():
    return sorted({k.split("/")[0] for k in STEP.files})


def case_desc(name):
    meta = json.loads(str(STEP[f"{name}/meta"]))
    return vs.VaeDesc(meta["batch"], meta["data_dim"], meta["latent_dim"], meta["enc_hidden"], meta["dec_hidden"],
                      sig_decoder=meta["sig_decoder"], tunable_decoder_var=meta["tunable"], epsilon0=meta["epsilon0"],
                      lr=meta["lr"])


# ------------------------------------------------------------------ oracle vs golden (CPU)
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_rng_matches_golden(impl):
    pub, der = KAT["published"], KAT["derived"]
    for v in pub["threefry2x32_kat"]:
        y0, y1 = jr.threefry2x32(U32(v["key"][0]), U32(v["key"][1]), np.array([v["ctr"][0]], U32), np.array([v["ctr"][1]], U32))
        assert [int(y0[0]), int(y1[0])] == v["out"]
    split = jr.split if impl == "numpy" else co.split
    normal = jr.normal if impl == "numpy" else co.normal
    bits = jr.random_bits if impl == "numpy" else co.random_bits
    assert np.asarray(split(jr.PRNGKey(0))).tolist() == pub["split_prngkey0"]
    assert np.float32(np.asarray(normal(jr.PRNGKey(0), (1,)))[0]) == np.float32(pub["normal_prngkey0"])
    assert np.float32(np.asarray(normal(jr.PRNGKey(42), (1,)))[0]) == np.float32(pub["normal_prngkey42"])
    assert np.float32(jr.uniform(jr.PRNGKey(0), ())) == np.float32(pub["uniform_prngkey0"])
    key = jr.PRNGKey(0)
    for i, want in enumerate(der["model_chain"]["consumed"]):
        rows = np.asarray(split(key))
        assert rows[0].tolist() == want                           # row 0 consumed
        key = rows[1]                                             # row 1 carried
        if str(i) in der["model_chain"]["carried_after"]:
            assert key.tolist() == der["model_chain"]["carried_after"][str(i)]
    lat = der["latent_step0"]
    b = bits(np.array(lat["key"], U32), lat["n"])
    assert b[:3].tolist() == lat["bits_0_3"] and b[650:652].tolist() == lat["bits_650_652"]
    z = np.asarray(normal(np.array(lat["key"], U32), tuple(lat["shape"])))
    np.testing.assert_allclose(z[0, :4], lat["z_row0_cols0_4"], rtol=lat["rtol"])
    np.testing.assert_allclose(z[50, :2], lat["z_row50_cols0_2"], rtol=lat["rtol"])
    for seed, v in der["dataset_chain_first_get_key"].items():
        if seed.startswith("_"):
            continue
        rows = np.asarray(split(jr.PRNGKey(int(seed))))
        assert rows[0].tolist() == v["carried"] and rows[1].tolist() == v["returned"]


@pytest.mark.parametrize("name", step_cases())
def test_oracle_step_matches_golden(name):
    """float32 NumPy oracle and the C port against the committed float64 vectors."""
    d = case_desc(name)
    P, x, z = STEP[f"{name}/params"], STEP[f"{name}/x"], STEP[f"{name}/z"]
    L = d.latent_dim
    loss, grad = vs.loss_and_grad(d, P, x, z[:, :L], z[:, L:], dtype=np.float32)
    gref = STEP[f"{name}/grad"]
    assert abs(loss - float(STEP[f"{name}/loss"])) <= 1e-5 * abs(float(STEP[f"{name}/loss"]))
    assert np.max(np.abs(grad - gref)) <= 1e-5 * np.max(np.abs(gref))
    closs, cgrad = co.loss_and_grad(co.make_desc(d), P, x, z)
    assert abs(closs - float(STEP[f"{name}/loss"])) <= 1e-5 * abs(float(STEP[f"{name}/loss"]))
    assert np.max(np.abs(cgrad - gref)) <= 1e-5 * np.max(np.abs(gref))


# ------------------------------------------------------------------ CUDA vs golden (GPU)
@pytest.mark.gpu
def test_cuda_rng_matches_golden():
    import torch
    from vae_training_b200 import ops
    pub, der = KAT["published"], KAT["derived"]
    assert ops.random_split(ops.key_tensor(jr.PRNGKey(0))).cpu().numpy().tolist() == pub["split_prngkey0"]
    for seed, want in ((0, pub["normal_prngkey0"]), (42, pub["normal_prngkey42"])):
        got = ops.random_normal(ops.key_tensor(jr.PRNGKey(seed)), (1,)).cpu().numpy()
        assert got.view(U32)[0] == np.float32(want).view(U32)
    key = ops.key_tensor(jr.PRNGKey(0))
    for want in der["model_chain"]["consumed"]:
        rows = ops.random_split(key)
        assert rows[0].cpu().numpy().tolist() == want
        key = rows[1].contiguous()
    lat = der["latent_step0"]
    k = ops.key_tensor(np.array(lat["key"], U32))
    b = ops.random_bits(k, lat["n"]).cpu().numpy()
    assert b[:3].tolist() == lat["bits_0_3"] and b[650:652].tolist() == lat["bits_650_652"]
    z = ops.random_normal(k, tuple(lat["shape"])).cpu().numpy()
    np.testing.assert_allclose(z[0, :4], lat["z_row0_cols0_4"], rtol=lat["rtol"])
    np.testing.assert_allclose(z[50, :2], lat["z_row50_cols0_2"], rtol=lat["rtol"])
    torch.cuda.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize("name", step_cases())
def test_cuda_step_matches_golden(name):
    import torch
    from vae_training_b200 import _lib, ops
    d = case_desc(name)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    P, x, z = dev(STEP[f"{name}/params"]), dev(STEP[f"{name}/x"]), dev(STEP[f"{name}/z"])
    cd = _lib.make_desc(d.batch, d.data_dim, d.latent_dim, list(d.enc_hidden), list(d.dec_hidden), sig_decoder=d.sig_decoder,
                        tunable_decoder_var=d.tunable_decoder_var, epsilon0=d.epsilon0, lr=d.lr)
    grad, loss = torch.zeros_like(P), torch.zeros(1, device="cuda")
    ws = ops.Workspace()
    ops.loss_and_grad(cd, P, x, z, d.batch, d.batch, loss, grad, ws)
    lref, gref = float(STEP[f"{name}/loss"]), STEP[f"{name}/grad"]
    assert abs(float(loss) - lref) <= 1e-5 * abs(lref)
    assert np.max(np.abs(grad.cpu().numpy() - gref)) <= 1e-5 * np.max(np.abs(gref))
    M, V = torch.zeros_like(P), torch.zeros_like(P)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.train_step(cd, P, M, V, step, x, z, loss, ws)
    pref = STEP[f"{name}/params_after"]
    # first Adam step moves every coordinate by ~lr; 1e-3 * lr + half an ulp of slack
    tol = 1e-3 * d.lr + 0.5 * np.spacing(np.abs(pref).astype(np.float32))
    frac_ok = np.mean(np.abs(P.cpu().numpy() - pref) <= tol)
    assert frac_ok >= 0.99, frac_ok
