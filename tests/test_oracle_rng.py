"""Pins oracle/jax_random.py against every published known answer available for the
path: Random123 threefry2x32 KATs, documented jax.random
values, and the derived key-chain values of Appendix A.5."""
import numpy as np
import pytest

from oracle import jax_random as jr

U32 = np.uint32


@pytest.mark.parametrize("key,ctr,out", [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6b200159, 0x99ba4efe)),
    ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
    ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0)),
])
def test_threefry_random123_kat(key, ctr, out):
    y0, y1 = jr.threefry2x32(U32(key[0]), U32(key[1]), np.array([ctr[0]], U32), np.array([ctr[1]], U32))
    assert (int(y0[0]), int(y1[0])) == out


def test_prngkey_and_split_documented_values():
    assert jr.PRNGKey(0).tolist() == [0, 0]
    assert jr.PRNGKey(69).tolist() == [0, 69]
    assert jr.PRNGKey((5 << 32) | 7).tolist() == [5, 7]
    assert jr.split(jr.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


def test_uniform_and_normal_documented_values():
    assert np.float32(jr.uniform(jr.PRNGKey(0), ())) == np.float32(0.41845703)
    assert np.float32(jr.normal(jr.PRNGKey(0), ())) == np.float32(-0.20584226)
    assert np.float32(jr.normal(jr.PRNGKey(42), ())) == np.float32(-0.18471177)


def test_odd_length_counts_are_padded():
    key = jr.PRNGKey(7)
    b5 = jr.random_bits(key, 5)
    # n=5 -> counters padded to [0,1,2,3,4,0]: block j uses (j, j+3), last pad counter 0
    y0, y1 = jr.threefry2x32(key[0], key[1], np.array([0, 1, 2], U32), np.array([3, 4, 0], U32))
    assert b5.tolist() == np.concatenate([y0, y1])[:5].tolist()


def test_model_chain_values():
    """model-side chain (row 0 consumed, row 1 carried)."""
    key = jr.PRNGKey(0)
    vae_key, key = jr.model_chain_next(key)                  # vae.py:56
    assert vae_key.tolist() == [4146024105, 967050713]
    assert key.tolist() == [2718843009, 1272950319]
    eval_key, key = jr.model_chain_next(key)                 # model.py:208
    assert eval_key.tolist() == [3746991216, 1902760697]
    assert key.tolist() == [1190051861, 3378399878]
    stats_key, key = jr.model_chain_next(key)                # model.py:154
    assert stats_key.tolist() == [2415289872, 3870638741]
    plot_key, key = jr.model_chain_next(key)                 # model.py:140
    assert plot_key.tolist() == [1908869515, 1172436361]
    lat0, key = jr.model_chain_next(key)                     # vae.py:125 step 0
    assert lat0.tolist() == [145021652, 2649282573]
    assert key.tolist() == [2088994074, 326295003]
    lat1, key = jr.model_chain_next(key)
    assert lat1.tolist() == [1409544366, 3423741852]
    bits = jr.random_bits(lat0, 1300)
    assert [hex(int(b)) for b in bits[:3]] == ["0xe36d7a59", "0xb4d702cd", "0xb383de42"]
    assert [hex(int(b)) for b in bits[650:652]] == ["0xdfefa6ee", "0x1537907"]
    z = jr.normal(lat0, (100, 13))
    np.testing.assert_allclose(z[0, :4], [1.2180066, 0.54291487, 0.52794409, 0.31873378], rtol=2e-7)
    np.testing.assert_allclose(z[50, :2], [1.1491385, -2.5635822], rtol=2e-7)


@pytest.mark.parametrize("seed,carried,returned", [
    (69, [3746304321, 1047182288], [3745723190, 1458479275]),
    (24, [3955744417, 3485677438], [2795415897, 2028976953]),
    (48, [3521067677, 3743923828], [2307453418, 1374000489]),
    (2, [2859854988, 796765312], [2425776485, 230565590]),
])
def test_dataset_chain_values(seed, carried, returned):
    """datasets.py:50-52: row 0 carried, row 1 returned."""
    ret, car = jr.dataset_chain_next(jr.PRNGKey(seed))
    assert car.tolist() == carried
    assert ret.tolist() == returned


def test_normal_statistics_and_tail_branch():
    z = jr.normal(jr.PRNGKey(123), (400000,))
    assert z.dtype == np.float32 and np.isfinite(z).all()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert (np.abs(z) > 2.93).mean() > 0.002            # tail polynomial exercised
    # erf_inv against the float64 definition (Giles' polynomial is a ~1e-6-accurate
    # approximation, not a correctly rounded one): central and tail branches
    from scipy.special import erfinv
    u = np.linspace(-0.9999999, 0.9999999, 200001).astype(np.float32)
    ref = erfinv(u.astype(np.float64))
    got = jr.erf_inv(u).astype(np.float64)
    nz = np.abs(ref) > 0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    central = -np.log1p(-u[nz].astype(np.float64) ** 2) < 5.0
    assert rel[central].max() <= 1e-6 and rel[~central].max() <= 1e-5


def test_neg_log1p_neg_accuracy():
    t = np.concatenate([np.linspace(0, 0.999999, 100001), 10.0 ** np.linspace(-30, -1, 1000)]).astype(np.float32)
    ref = -np.log1p(-t.astype(np.float64))
    got = jr.neg_log1p_neg(t).astype(np.float64)
    ulp = np.spacing(np.maximum(np.abs(ref), 1e-45).astype(np.float32)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ulp) <= 1.5
