"""Host-side logic that needs no GPU: the run.py flag surface, sweep-script parsing, the
LPT partition, the checkpoint layout, parameter init, event schedule arithmetic."""
import os

import numpy as np
import pytest

from vae_training_b200 import checkpoint, init, run, sweep
from vae_training_b200.packed import RunSpec

SIGMOID_LINE = ('python run.py sigmoid_dd3_pd3_ld_6_eps-3 --dataset sigmoid --encoder_layer_sizes "" '
                '--layer_sizes "" -ow --latent_dim 6  --padding_dim 3 -dd 3 --num_batches 150000 --epsilon -3 -tdv')
SPHERE_LINE = ('python run.py sphere_dd5_pd16_ld_16_eps-3 --dataset sphere --encoder_layer_sizes "200|200|200" '
               '--layer_sizes "200|200|200" -ow --latent_dim 16  --padding_dim 16 -dd 5 --num_batches 150000 '
               '--epsilon -3 -tdv -ds 24')
LINEAR_LINE = ('python run.py vae3linear_gaussian_12dim2 --dataset linear_gaussian --encoder_layer_sizes "" '
               '--layer_sizes "" -ow --latent_dim 20  --padding_dim 9 -dd 3 --num_batches 100000  --epsilon -1 -tdv '
               '-ds 2  -lr 1e-3')


def test_flag_surface_defaults_match_reference():
    a = run.parse_arguments(["x"])
    assert (a.num_batches, a.batch_size, a.learning_rate, a.dataset_seed) == (15000, 100, 0.0001, 69)
    assert (a.latent_dimension, a.layer_sizes, a.encoder_layer_sizes) == (100, "512|512", "512|512")
    assert (a.epsilon, a.tunable_decoder_var, a.dataset_dimension, a.dataset_intrinsic_dimension) == (0.0, False, 3, 3)
    assert (a.model, a.latent_distribution, a.tqdm) == ("VAE", "gaussian", True)


def test_sweep_script_parsing(tmp_path):
    p = tmp_path / "expts.sh"
    p.write_text("\n".join([SIGMOID_LINE, SPHERE_LINE, "", LINEAR_LINE, "echo done"]) + "\n")
    runs = sweep.parse_script(str(p))
    assert [r.name for r in runs] == ["sigmoid_dd3_pd3_ld_6_eps-3", "sphere_dd5_pd16_ld_16_eps-3", "vae3linear_gaussian_12dim2"]
    s, sp, li = runs
    assert (s.dataset, s.latent_dimension, s.padding_dim, s.dataset_dimension, s.epsilon, s.layer_sizes) == ("sigmoid", 6, 3, 3, -3.0, "")
    assert (sp.layer_sizes, sp.dataset_seed, sp.num_batches) == ("200|200|200", 24, 150000)
    assert (li.learning_rate, li.dataset_seed, li.epsilon, li.tunable_decoder_var) == (1e-3, 2, -1.0, True)


def test_lpt_partition_balances_and_covers():
    costs = [104.0] * 15 + [1.0] * 39
    parts = sweep.lpt_partition(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(54))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 104.0
    assert sweep.lpt_partition([3.0, 1.0], 1) == [[0, 1]]


def test_latency_partition_spreads_the_hidden_layer_chains():
    """The 54 runs over 8 GPUs: the 15 sphere chains (150 000 serial steps each -- the wall time) go two per GPU, the 39
    single-Dense runs are spread beside them; every run lands on exactly one rank for every world size."""
    runs = sweep.builtin_runs("all")
    for world in (1, 2, 3, 4, 8):
        parts = sweep.latency_partition(runs, world)
        assert sorted(sum(parts, [])) == list(range(len(runs)))
        sph = [sum(1 for i in p if runs[i].dataset == "sphere") for p in parts]
        assert max(sph) - min(sph) <= 1
        small = [len(p) - s for p, s in zip(parts, sph)]
        assert max(small) - min(small) <= 3
    parts = sweep.latency_partition(runs, 8)
    assert max(sweep.estimated_wall_seconds(runs, p) for p in parts) <= 150000 * 150e-6 + 1e-9     # one 16-CTA chain latency
    assert sweep.chain_latency_us(1) < sweep.chain_latency_us(15) < sweep.chain_latency_us(33) < sweep.chain_latency_us(148)


def test_runspec_shapes_match_survey_appendix_c():
    s = RunSpec("sigmoid", 7, 20, 24, tunable_decoder_var=True)
    assert s.data_dim == 28 and s.flops_per_step() == 1075200 and s.normals_per_step() == 5900
    h = [200, 200, 200]
    s = RunSpec("sphere", 3, 3, 6, h, h, tunable_decoder_var=True)
    assert s.data_dim == 6 and s.flops_per_step() == 98640000 and s.normals_per_step() == 1500
    s = RunSpec("linear_gaussian", 3, 9, 20, tunable_decoder_var=True)
    assert s.data_dim == 12 and s.flops_per_step() == 240000 and s.normals_per_step() == 3500


def test_init_params_distribution_and_layout():
    h = [200, 200, 200]
    spec = RunSpec("sphere", 5, 16, 16, h, h, tunable_decoder_var=True)
    p = init.init_params(spec, np.array([1, 2], np.uint32))
    shapes = init.layer_shapes(spec)
    assert p.size == sum(int(np.prod(s)) for _, s in shapes) == 176054
    tree = checkpoint.unflatten(spec, p)
    k = tree["Encoder/FC1/kernel"]
    assert k.shape == (200, 200) and abs(k.std() - np.sqrt(1 / 200)) < 0.003
    assert np.abs(k).max() <= 2.0 * np.sqrt(1 / 200) / 0.87962566 + 1e-6          # truncated at 2 sigma
    assert (tree["Decoder/FC3/bias"] == 0).all() and (tree["epsilon_p"] == 1).all() and (tree["epsilon"] == 1).all()
    assert [n for n, _ in shapes][:2] == ["Encoder/FC0/kernel", "Encoder/FC0/bias"]
    assert [n for n, _ in shapes][-2:] == ["epsilon_p", "epsilon"]
    # deterministic in the vae_key
    np.testing.assert_array_equal(p, init.init_params(spec, np.array([1, 2], np.uint32)))


def test_init_rng_known_answers():
    """The host threefry of init.py against the Random123 known answers and the documented split(PRNGKey(0))."""
    import json
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "jax_random_kat.json")))
    for v in kat["published"]["threefry2x32_kat"]:
        with np.errstate(over="ignore"):
            y0, y1 = init._threefry2x32(np.uint32(v["key"][0]), np.uint32(v["key"][1]),
                                        np.array([v["ctr"][0]], np.uint32), np.array([v["ctr"][1]], np.uint32))
        assert [int(y0[0]), int(y1[0])] == v["out"]
    # fold_in(key, d) = threefry(key, [0, d]) in jax's array layout (lane 0 counter 0, lane 1 counter d)
    key = np.array([0x243F6A88, 0x85A308D3], np.uint32)
    with np.errstate(over="ignore"):
        y0, y1 = init._threefry2x32(key[0], key[1], np.array([0], np.uint32), np.array([77], np.uint32))
    np.testing.assert_array_equal(init.fold_in(key, 77), [y0[0], y1[0]])
    # sha1("kernel")[:4] big endian (legacy flax _fold_in_str)
    import hashlib
    assert int.from_bytes(hashlib.sha1(b"kernel").digest()[:4], "big") == int(hashlib.sha1(b"kernel").hexdigest()[:8], 16)
    np.testing.assert_array_equal(init.leaf_key(key, "Encoder/FC0/kernel"),
                                  init.fold_in_str(init.fold_in_str(init.fold_in_str(key, "Encoder"), "FC0"), "kernel"))


def test_truncated_normal_definition():
    """jax.random.truncated_normal(key, -2, 2): inside the open interval, the distribution of a unit normal cut at 2 sigma."""
    x = init.truncated_normal(np.array([5, 6], np.uint32), -2.0, 2.0, (300, 401))
    assert x.dtype == np.float32 and x.shape == (300, 401)
    assert x.min() > -2.0 and x.max() < 2.0
    assert abs(float(x.mean())) < 0.01 and abs(float(x.std()) - 0.87962566) < 0.005
    # odd element counts use jax's padded counter layout; a prefix of a longer draw is a different stream
    y = init.truncated_normal(np.array([5, 6], np.uint32), -2.0, 2.0, (7,))
    assert y.shape == (7,) and not np.array_equal(y, x.reshape(-1)[:7])


REF_INIT_FIXTURES = ["ref_linear_dd3_pd9_ld20", "ref_linear_noise_dn", "ref_sigmoid_dd3_pd3_ld6",
                     "ref_sigmoid_dd5_pd5_ld10_seed24", "ref_sphere_dd3_pd3_ld6", "ref_sphere_small_fixedvar"]


@pytest.mark.parametrize("name", REF_INIT_FIXTURES)
def test_init_params_reproduce_the_reference_init(name):
    """`init_params` of the fixtures came out of the reference's OWN `VAE.init_by_shape(vae_key, ...)` (vae.py:56-60,
    networks.py:36,63-77) executed on tests/refshim, whose flax.nn restates legacy flax's per-name key folding and
    computes the truncated normal in float64.  The product's float32 init must land on the same values: same leaves in
    the same order, same names folded, same key chain (vae_key = row 0 of split(PRNGKey(0)), model.py:29), same
    distribution -- to float32 round-off (<= 5e-7 absolute on values of magnitude <= 1.3)."""
    import dataclasses
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"), allow_pickle=False)
    a = run.parse_arguments([str(x) for x in f["argv"]])
    spec = dataclasses.replace(sweep._spec_of(a), tunable_decoder_var=a.tunable_decoder_var)
    vae_key = init._threefry_counts(np.zeros(2, np.uint32), np.arange(4, dtype=np.uint32)).reshape(2, 2)[0]
    got = init.init_params(spec, vae_key)
    assert got.dtype == np.float32 and got.shape == f["init_params"].shape
    assert np.abs(got.astype(np.float64) - f["init_params"]).max() <= 5e-7
    # and a different module key gives different kernels (the fixture is not matched by accident)
    other = init.init_params(spec, vae_key ^ np.uint32(1))
    assert np.abs(other.astype(np.float64) - f["init_params"]).max() > 1e-2


def test_checkpoint_roundtrip(tmp_path):
    spec = RunSpec("sigmoid", 3, 3, 6, tunable_decoder_var=True)
    rng = np.random.default_rng(0)
    p, m, v = (rng.standard_normal(153).astype(np.float32) for _ in range(3))
    sd = checkpoint.to_state_dict(spec, p, m, v, 17)
    assert set(sd) == {"target", "state"} and sd["state"]["step"] == 17
    assert set(sd["target"]["params"]) == {"Encoder", "Decoder", "SigDecoder", "epsilon_p", "epsilon"}
    assert sd["target"]["params"]["Encoder"]["FC0"]["kernel"].shape == (7, 6)
    assert set(sd["state"]["param_states"]["Decoder"]["FC0"]["bias"]) == {"grad_ema", "grad_sq_ema"}
    fn = str(tmp_path / "model.pkl")
    checkpoint.save_model(fn, spec, p, m, v, 17)
    p2, m2, v2, st = checkpoint.load_model(fn, spec)
    np.testing.assert_array_equal(p, p2)
    np.testing.assert_array_equal(m, m2)
    np.testing.assert_array_equal(v, v2)
    assert st == 17


def test_make_output_dir_rules(tmp_path):
    a = run.parse_arguments(["exp1", "-ow"])
    d = run.make_output_dir("exp1", False, a, data_dir=str(tmp_path))
    assert os.path.exists(os.path.join(d, "args.json"))
    with pytest.raises(ValueError):
        run.make_output_dir("exp1", False, a, data_dir=str(tmp_path))
    open(os.path.join(d, "junk"), "w").close()
    run.make_output_dir("exp1", True, a, data_dir=str(tmp_path))
    assert sorted(os.listdir(d)) == ["args.json"]


def test_builtin_sweeps_match_reference_scripts():
    """sweeps.py regenerates the three .sh files: 21 + 18 + 15 runs with the same flags.  When
    the reference tree is present (this container) the regenerated argument namespaces are
    compared with the parsed scripts run by run."""
    from vae_training_b200 import sweeps
    assert (len(sweeps.linear_runs()), len(sweeps.sigmoid_runs()), len(sweeps.sphere_runs())) == (21, 18, 15)
    runs = sweep.builtin_runs("all")
    assert len(runs) == 54 and sum(r.num_batches for r in runs) == 7_050_000
    ref_dir = "/root/reference"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not available")
    key = lambda a: (a.dataset, a.dataset_dimension, a.padding_dim, a.latent_dimension, a.dataset_seed)
    fields = ("name", "dataset", "dataset_dimension", "padding_dim", "latent_dimension", "dataset_seed", "num_batches",
              "learning_rate", "epsilon", "tunable_decoder_var", "layer_sizes", "encoder_layer_sizes", "batch_size",
              "overwrite")
    for which, fn in (("linear", "seed_linpadding_expts.sh"), ("sigmoid", "sigmoid_vae_padding_expts.sh"),
                      ("sphere", "sphere_vae_padding_expts.sh")):
        ref = {key(a): a for a in sweep.parse_script(os.path.join(ref_dir, fn))}
        mine = {key(a): a for a in sweep.builtin_runs(which)}
        assert set(ref) == set(mine), which
        for k in ref:
            for f in fields:
                assert getattr(ref[k], f) == getattr(mine[k], f), (which, k, f)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON
    line with the contract's keys; here on a tiny bounded sample so the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "300", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "train-steps/s" and d["higher_is_better"] is True
    assert d["metric"] == "vae_train_steps_per_sec_summed_over_packed_models"
    assert d["steps"] == 300 and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
