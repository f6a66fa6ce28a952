"""GPU tests of the host-side mirror of the reference interface (datasets.py, vae.py,
run.py, sweep.py): the full train() schedule with eval / plot events against the oracle
trainer, packed sweeps, output files."""
import os

import numpy as np
import pytest
import torch

from oracle import datasets as ods
from oracle import trainer as otr
from oracle import vae_step as vs

pytestmark = pytest.mark.gpu


def build_pair(name, dd, pd, L, hidden, N, seed=69, lr=1e-4, eps0=-3.0, n_print=8, n_plot=16, eval_batch=50):
    from vae_training_b200 import datasets, vae
    cls = {"sphere": datasets.SphereDataset, "sigmoid": datasets.SigmoidDataset}
    if name == "linear_gaussian":
        ds = datasets.LinearGaussianDataset(seed, dimension=dd, intrinsic_dimension=3, padding_dimension=pd)
    else:
        ds = cls[name](seed, dimension=dd, padding_dimension=pd)
    ods_ = ods.make_dataset(name, seed, dd, pd)
    if ods_.A is not None:
        np.testing.assert_array_equal(ds.A.cpu().numpy().reshape(ods_.A.shape), ods_.A)
    hid = [int(v) for v in hidden.split("|")] if hidden else []
    d = vs.VaeDesc(100, ods_.ndim, L, hid, hid, sig_decoder=(name == "sigmoid"), epsilon0=eps0, lr=lr)
    flat = vs.init_params(d, 0)
    m = vae.VAEModel(None, N, 1, 100, lr, hidden, hidden, None, None, eps0, False, ds, L, tunable_decoder_var=True,
                     dataset_name=name, initial_params=flat, verbose=False)
    m.n_print, m.n_plot, m.print_batch_size, m.plot_batch_size = n_print, n_plot, eval_batch, eval_batch
    tr = otr.OracleTrainer(d, ods_, flat, n_print=n_print, n_plot=n_plot, eval_batch=eval_batch, dtype=np.float64)
    return m, tr, ods_


def compare(m, tr, ods_, N):
    got = m.vae_losses.astype(np.float64)
    ref = np.array(tr.vae_losses, np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-5
    np.testing.assert_array_equal(m.key.cpu().numpy(), tr.key)
    np.testing.assert_array_equal(m.dataset.key.cpu().numpy(), ods_.key)
    assert len(m.var_enc) == len(tr.var_enc)
    for a, b in zip(m.var_enc, tr.var_enc):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=1e-4, atol=1e-6)
    for a, b in zip(m.var_dec, tr.var_dec):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=1e-4, atol=1e-6)
    assert set(m.stats) == set(tr.stats)
    for k in tr.stats:
        np.testing.assert_allclose(np.array(m.stats[k], np.float64), np.array(tr.stats[k], np.float64), rtol=2e-4, atol=1e-6)
    assert int(m.step) == N


@pytest.mark.parametrize("name,dd,pd,L,hidden", [("sigmoid", 3, 3, 6, ""), ("linear_gaussian", 3, 9, 20, ""),
                                                 ("sphere", 3, 3, 6, "24|16")])
def test_train_schedule_matches_oracle(name, dd, pd, L, hidden):
    """model.py:207-222 end to end (single-CTA fused kernel for the single-Dense shapes, cluster
    kernel for the hidden-layer one): interleaved train/eval losses, learned variances, stats, key chains."""
    N = 30
    m, tr, ods_ = build_pair(name, dd, pd, L, hidden, N, seed=2 if name == "linear_gaussian" else 69)
    assert m.fused_class == (1 if hidden == "" else 2)
    m.train_distribution(save=False)
    tr.train(N)
    compare(m, tr, ods_, N)


def test_dataset_noise_takes_the_per_op_path():
    """`-dn` / var_added > 0 (datasets.py:189-192) draws a second key and adds noise to every batch:
    the fused kernels do not model it, so such a run must fall back to the per-op step and still
    follow the oracle's schedule (loss, key chains, stats)."""
    from vae_training_b200 import datasets, vae
    N, seed, dd, pd, L = 12, 2, 3, 9, 20
    ds = datasets.LinearGaussianDataset(seed, dimension=dd, intrinsic_dimension=3, padding_dimension=pd, var_added=0.25)
    ods_ = ods.make_dataset("linear_gaussian", seed, dd, pd, var_added=0.25)
    d = vs.VaeDesc(100, ods_.ndim, L, [], [], epsilon0=-1.0, lr=1e-3)
    flat = vs.init_params(d, 0)
    m = vae.VAEModel(None, N, 1, 100, 1e-3, "", "", None, None, -1.0, False, ds, L, tunable_decoder_var=True,
                     dataset_name="linear_gaussian", initial_params=flat, verbose=False)
    assert not m.fused and m.fused_class == 0
    m.n_print, m.n_plot, m.print_batch_size, m.plot_batch_size = 4, 8, 50, 50
    tr = otr.OracleTrainer(d, ods_, flat, n_print=4, n_plot=8, eval_batch=50, dtype=np.float64)
    m.train_distribution(save=False)
    tr.train(N)
    compare(m, tr, ods_, N)


def test_train_many_equals_individual_runs():
    from vae_training_b200 import vae
    cfgs = [("sigmoid", 3, 3, 6, "", 30), ("sigmoid", 5, 5, 10, "", 22), ("linear_gaussian", 6, 6, 20, "", 17)]
    packed = [build_pair(*c[:5], c[5]) for c in cfgs]
    vae.train_many([p[0] for p in packed], save=False)
    for (m, tr, ods_), c in zip(packed, cfgs):
        tr.train(c[5])
        compare(m, tr, ods_, c[5])


def test_sweep_outputs(tmp_path, monkeypatch):
    from vae_training_b200 import run, sweep, vae
    monkeypatch.setattr(run, "DATA_DIR", str(tmp_path))
    monkeypatch.setattr(vae.VAEModel, "n_print", 10)
    monkeypatch.setattr(vae.VAEModel, "n_plot", 20)
    script = tmp_path / "expts.sh"
    script.write_text(
        'python run.py s1 --dataset sigmoid --encoder_layer_sizes "" --layer_sizes "" -ow --latent_dim 6 --padding_dim 3 -dd 3 --num_batches 25 --epsilon -3 -tdv\n'
        'python run.py l1 --dataset linear_gaussian --encoder_layer_sizes "" --layer_sizes "" -ow --latent_dim 20 --padding_dim 9 -dd 3 --num_batches 25 --epsilon -1 -tdv -ds 2 -lr 1e-3\n')
    models = sweep.run_sweep(sweep.parse_script(str(script)))
    assert len(models) == 2
    for name in ("s1", "l1"):
        d = tmp_path / name
        assert sorted(os.listdir(d)) == ["args.json", "losses.npz", "model.pkl"]
        z = np.load(d / "losses.npz", allow_pickle=True)
        for k in ("VAE Loss", "Decoder Variance", "Encoder Variance", "EigenValues", "Average Log Likelihood",
                  "Correlation Ratio", "KL divergence", "mse"):
            assert k in z.files, k
        assert z["VAE Loss"].shape == (25 + 3,)          # N train losses + one per stats event (vae.py:130,136)
        assert z["Encoder Variance"].shape[0] == 3 and np.isfinite(z["VAE Loss"]).all()
    import pickle
    sd = pickle.load(open(tmp_path / "s1" / "model.pkl", "rb"))
    assert sd["state"]["step"] == 25 and "SigDecoder" in sd["target"]["params"]


def test_dataset_classes_follow_reference_interface():
    from vae_training_b200 import datasets
    ds = datasets.SphereDataset(69, dimension=3, padding_dimension=3)
    assert ds.shape == (6,) and ds.dimension == 6 and ds.is_epochs is False
    k0 = ds.key.clone()
    b, lat = ds.get_batch(100, return_latents=True)
    assert lat is None and b.shape == (100, 6) and not torch.equal(k0, ds.key)
    sc = ds.score_batch(b)
    assert set(sc) == {"Sphere Error", "Padding Error"} and sc["Sphere Error"] < 1e-10 and sc["Padding Error"] == 0
    ref = ods.SphereDataset(69, 3, 3).get_batch(100)
    np.testing.assert_array_equal(b.cpu().numpy(), ref)
    noisy = datasets.LinearGaussianDataset(2, dimension=3, intrinsic_dimension=3, padding_dimension=2, var_added=0.25)
    ref2 = ods.LinearGaussianDataset(2, 3, 3, 2, var_added=0.25)
    np.testing.assert_allclose(noisy.get_batch(50).cpu().numpy(), ref2.get_batch(50), rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(noisy.key.cpu().numpy(), ref2.key)


def test_bench_b200_arm_contract():
    """`bench.py` (GPU arm) prints ONE JSON line carrying every key of the measurement contract;
    short run, CPU baseline included on a 1-second sample."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "20", "--warmup", "5", "--cpu-seconds", "1",
                          "--dp-steps", "1"], capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["steps"] == 20 and d["warmup"] == 5 and d["n_gpus"] == 1 and d["value"] > 0
    # one persistent launch per repeat of the K-step unit; the window is >= 50 ms of back-to-back launches
    assert d["gpu_launches"] == d["timing"]["repeats"] >= 3 and d["timing"]["window_ms"] >= 45.0
    assert abs(d["ms_per_step"] - d["timing"]["launch_ms_median"] / 20) < 1e-12
    # the other BASELINE.json configs ride in the same line, each with its own roofline
    assert set(d["sub_results"]) == {"single_chain", "sphere_pack", "sphere_single_chain", "sweep_mix_54_runs", "cfg5_sphere_dp"}
    for name, sub in d["sub_results"].items():
        assert sub["value"] > 0 and sub["ms_per_step"] > 0, name
        assert set(sub["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}, name
    assert d["sub_results"]["cfg5_sphere_dp"]["roofline"]["bound"] == "tensor"
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_all_54_sweep_runs_short_match_oracle(monkeypatch):
    """BASELINE configs[3]: every run of the three sweep scripts (54 runs: 13 single-Dense shapes, 5
    hidden-layer shapes, 3 dataset seeds each), built through the run.py flag surface and trained
    packed by `train_many` -- here for 11 steps with the event cadence shrunk to 5 / 10 -- must follow
    the oracle trainer run by run: interleaved train / eval losses rel 1e-5, key chains bit-exact,
    learned variances and stats."""
    from vae_training_b200 import run, sweep, vae
    N = 11
    monkeypatch.setattr(vae.VAEModel, "n_print", 5)
    monkeypatch.setattr(vae.VAEModel, "n_plot", 10)
    monkeypatch.setattr(vae.VAEModel, "print_batch_size", 60)
    monkeypatch.setattr(vae.VAEModel, "plot_batch_size", 60)
    all_args = sweep.builtin_runs("all", N)
    assert len(all_args) == 54
    models, trainers = [], []
    for a in all_args:
        ds = run.get_dataset(a.dataset, a.dataset_seed, a.padding_dim, a.batch_size, a)
        m = run.get_model(a, ds, None, verbose=False)
        ods_ = ods.make_dataset(a.dataset, a.dataset_seed, a.dataset_dimension, a.padding_dim)
        hid = [int(v) for v in a.layer_sizes.split("|")] if a.layer_sizes else []
        d = vs.VaeDesc(a.batch_size, ods_.ndim, a.latent_dimension, hid, hid, sig_decoder=(a.dataset == "sigmoid"),
                       epsilon0=a.epsilon, lr=a.learning_rate)
        flat = m.params.cpu().numpy().copy()
        assert flat.shape == (d.n_params,)
        tr = otr.OracleTrainer(d, ods_, flat, n_print=5, n_plot=10, eval_batch=60, dtype=np.float64)
        models.append(m)
        trainers.append((tr, ods_))
    assert sorted({m.fused_class for m in models}) == [1, 2]
    vae.train_many(models, save=False)
    for m, (tr, ods_) in zip(models, trainers):
        tr.train(N)
        compare(m, tr, ods_, N)


def test_gaussian_dataset_batches_and_scores():
    """datasets.py:101-160 (constructed directly; run.py never selects it): (size, dim + padding) batches from two
    key-chain draws (normal part, N(0, noise_level) padding) bit-identical to the jax.random restatement, and the
    reference's three score entries."""
    from oracle import jax_random as jr
    from vae_training_b200 import datasets
    ds = datasets.GaussianDataset(5, dimension=3, padding_dimension=2, noise_level=0.04)
    key = jr.PRNGKey(5)
    key, k1 = jr.split(key)
    key, k2 = jr.split(key)
    ref = np.concatenate([jr.normal(k1, (64, 3)), jr.normal(k2, (64, 2)) * np.float32(np.sqrt(np.float32(0.04)))], axis=1)
    got = ds.get_batch(64)
    np.testing.assert_array_equal(got.cpu().numpy().view(np.uint32), ref.astype(np.float32).view(np.uint32))
    np.testing.assert_array_equal(ds.key.cpu().numpy(), key)
    sc = ds.score_batch(got)
    assert set(sc) == {"Squared Norm of padding dimensions", "ground truth eigenvalue", "learnt eigenvalue"}
    assert abs(float(sc["Squared Norm of padding dimensions"]) - float((ref[:, 3:] ** 2).sum(1).mean())) < 1e-6
    assert sc["learnt eigenvalue"].shape == (5,)
    z = datasets.GaussianDataset(5, dimension=3, padding_dimension=2, noise_level=0.0).get_batch(8)
    assert float(z[:, 3:].abs().max()) == 0.0


def test_overlapped_pack_matches_single_pack():
    """packed.OverlappedPack (the end-to-end call of bench.py: sub-packs on their own streams, copies under compute,
    CUDA-graph replay) against one PackedRuns over the same runs: first-round losses and final host state bit-identical
    for eager, captured and pipelined-rounds execution; R pipelined rounds == R consecutive train_host calls."""
    import bench
    from vae_training_b200 import init
    from vae_training_b200.packed import OverlappedPack, PackedRuns
    K, R = 6, 3
    specs = bench.make_specs("sigmoid_dd3_pd3_ld6", 24)
    states = [init.initial_state(s) for s in specs]
    ref = PackedRuns(specs, max_chunk=K)
    ref_states = [init.initial_state(s) for s in specs]
    for _ in range(R):
        ref.train_host(ref_states, K)
    ref_first = np.stack([st.losses[0] for st in ref_states])
    ref_last = np.stack([st.losses[-1] for st in ref_states])
    for mode in ("joined", "rounds"):
        op = OverlappedPack(specs, max_chunk=K, parts=3)
        assert [hi - lo for lo, hi in op.bounds] == [8, 8, 8]
        op.prepare()
        for attempt in range(3):                 # eager, capture + replay, replay
            op.stage(states)
            if mode == "joined":
                op.train_host_all(K)
                torch.cuda.synchronize()
                np.testing.assert_array_equal(op.host_losses(K), ref_first)
            else:
                op.train_host_rounds(K, R)
                torch.cuda.synchronize()
                np.testing.assert_array_equal(op.host_losses(K), ref_last)
                out = [init.initial_state(s) for s in specs]
                for (lo, hi), p in zip(op.bounds, op.packs):
                    p.unpack(out[lo:hi], K)
                for a, b in zip(out, ref_states):
                    np.testing.assert_array_equal(a.params, b.params)
                    np.testing.assert_array_equal(a.model_key, b.model_key)
                    np.testing.assert_array_equal(a.dataset_key, b.dataset_key)
                    assert a.step == b.step == K * R
