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
