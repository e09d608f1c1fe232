"""GPU tests of the SURVEY.md section 8(f) rows ("next"): the native iterative-ladder estimator against the reference's
estimator and ladder (golden ladder_*.npz), numeric parity THROUGH the batched sweep drivers against the reference's
recorded data/ curve points, the reference's data/average_seeds.py on the drivers' output, and
MCMCSimulation_GPU.benchmark_performance."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests._util import load_golden, golden_names
from tests.test_gpu_parity import _shared_stream_se

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def _ladder_target(name):
    import rwm_pt_pytorch_b200.target_distributions as td
    if name == "ladder_rough_carpet_pm15_d20":
        return td.RoughCarpetDistributionTorch(20, device="cuda", mode_centers=[-15.0, 0.0, 15.0])
    return td.ThreeMixtureDistributionTorch(30, device="cuda", mode_centers=[[-15.0] + [0.0] * 29, [0.0] * 30, [15.0] + [0.0] * 29],
                                            mode_weights=[1 / 3, 1 / 3, 1 / 3])


def _pt(target, **kw):
    from rwm_pt_pytorch_b200.algorithms import ParallelTemperingRWM_GPU_Optimized as PT
    return PT(target.dim, 2.38 ** 2 / target.dim, target, device="cuda", swap_every=10, store="none", **kw)


@pytest.mark.parametrize("name", golden_names("ladder_"))
def test_native_swap_probability_estimator_matches_reference(name):
    """rwmpt_swap_prob_estimate (one kernel: Philox tempered samples at both temperatures, both log-densities, fp64 sum)
    against the reference's estimator (pt_rwm_gpu_optimized.py:356-368 on torch-CPU, 4e5 samples per pair) on every
    (beta, beta*) pair of the fixture: within 4.5 combined standard errors."""
    _cuda()
    _, g = load_golden(name)
    t = _ladder_target(name)
    algo = _pt(t, beta_ladder=[1.0, 0.5])        # any ladder: only the estimator is used here
    assert t.family_id in algo._NATIVE_LADDER_FAMILIES
    torch.manual_seed(3)
    N = 1_000_000
    for bc, bs, ref, se in zip(g["pair_beta"], g["pair_beta_star"], g["pair_estimate"], g["pair_se"]):
        est = algo._estimate_swap_probability(float(bc), float(bs), N)
        se_mine = np.sqrt(max(est * (1 - est), 1e-4) / N)     # variance of a [0,1] variable <= p(1-p)
        assert abs(est - ref) <= 4.5 * np.hypot(se, se_mine) + 1e-4, (float(bc), float(bs), est, float(ref), float(se))
    # the eager path (draw_samples_torch + CUDA log-density) and the native kernel estimate the same number
    bc, bs = float(g["pair_beta"][0]), float(g["pair_beta_star"][0])
    native = algo._estimate_swap_probability(bc, bs, N)
    xs, xc = t.draw_samples_torch(400_000, bs), t.draw_samples_torch(400_000, bc)
    eager = torch.mean(torch.exp(torch.clamp_max((bc - bs) * (t.log_density(xs) - t.log_density(xc)), 0.0))).item()
    assert abs(native - eager) <= 4.5 * np.hypot(0.5 / np.sqrt(N), 0.5 / np.sqrt(400_000))


@pytest.mark.parametrize("name", golden_names("ladder_"))
def test_iterative_ladder_matches_reference_rung_by_rung(name):
    """The ladder the drop-in builds with the native estimator against the reference's own ladder for the same target and
    settings: same number of rungs (+-1), every rung within 12 % (the reference accepts a rung when its 2e4-sample estimate
    is within 0.005 of the target, so its own rungs carry that much noise), and at the REFERENCE's rungs the native
    estimator returns the target rate within tolerance + Monte-Carlo error."""
    _cuda()
    _, g = load_golden(name)
    t = _ladder_target(name)
    torch.manual_seed(11)
    algo = _pt(t, iterative_temp_spacing=True, swap_acceptance_rate=float(g["target_rate"]), N_samples_swap_est=int(g["n_est"]))
    mine, ref = np.asarray(algo.beta_ladder), g["ladder"]
    assert algo.get_name().endswith("ITERATIVE_LADDER")
    assert abs(len(mine) - len(ref)) <= 1, (mine.tolist(), ref.tolist())
    m = min(len(mine), len(ref)) - 1
    np.testing.assert_allclose(mine[:m], ref[:m], rtol=0.12)
    assert mine[0] == 1.0 and abs(mine[-1] - 0.01) < 1e-9 and np.all(np.diff(mine) < 0)
    N = 2_000_000
    se_ref = np.sqrt(0.25 / float(g["n_est"]))
    for k in range(int(g["n_ladder_pairs"]) - 1):
        est = algo._estimate_swap_probability(float(ref[k]), float(ref[k + 1]), N)
        assert abs(est - float(g["target_rate"])) <= float(g["tolerance"]) + 4 * np.hypot(se_ref, 0.5 / np.sqrt(N)), (k, est)


# reference data/ points: (file pattern facts) target, var_max of the 40-value sweep, index -> (x, n_seeds, acc mean, acc sd,
# esjd mean, esjd sd) from data/<target>_Normal_RWM_GPU_dim20_1000000iters_seed*.json
RWM_DATA_POINTS = {
    "EvenRosenbrock": (0.6, 24, {5: (0.085641, 0.72332, 0.01442, 0.005213, 0.000108), 10: (0.161282, 0.46323, 0.03758, 0.011520, 0.000994),
                                 19: (0.297436, 0.18487, 0.02987, 0.014801, 0.002517)}),
    "FullRosenbrock": (1.3, 20, {5: (0.175385, 0.73417, 0.00311, 0.022198, 0.000097), 10: (0.340769, 0.51297, 0.00485, 0.057184, 0.000578),
                                 19: (0.638462, 0.23247, 0.00288, 0.085917, 0.001110)}),
    "NealFunnel": (6.6, 30, {5: (0.854872, 0.54140, 0.03218, 0.385569, 0.023266), 10: (1.699744, 0.35467, 0.06318, 0.981295, 0.176041),
                             19: (3.220513, 0.20530, 0.03783, 1.995305, 0.375152)}),
}


@pytest.mark.parametrize("target", sorted(RWM_DATA_POINTS))
def test_rwm_study_driver_matches_reference_data_points(target, tmp_path):
    """run_rwm_study (the 40-scale sweep of experiment_RWM_GPU.py:165-301 as ONE launch, 64 chains per scale, the
    reference's 1e6 iterations) against three recorded points of the reference's sweep for the same target: acceptance
    within 3 standard errors, ESJD within 2 % (+ 3 s.e.); the recorded seeds share one random stream (_shared_stream_se)."""
    _cuda()
    from rwm_pt_pytorch_b200.experiments import run_rwm_study
    var_max, n_seeds, pts = RWM_DATA_POINTS[target]
    out = run_rwm_study(20, target, num_iters=1_000_000, var_max=var_max, seed=3, burn_in=1000, chains_per_value=64,
                        out_dir=str(tmp_path))
    assert os.path.exists(out["filename"]) and len(out["acceptance_rates"]) == 40
    cpv = out["chains_per_value"]
    for idx, (x, acc_ref, acc_sd, esjd_ref, esjd_sd) in pts.items():
        assert abs(out["scale_param_range"][idx] - x) < 1e-5
        acc, esjd = out["acceptance_rates"][idx], out["expected_squared_jump_distances"][idx]
        acc_sd_ind, esjd_sd_ind = out["acceptance_rate_se"][idx] * np.sqrt(cpv), out["esjd_se"][idx] * np.sqrt(cpv)
        acc_err = np.hypot(out["acceptance_rate_se"][idx], _shared_stream_se(acc_sd, n_seeds, acc_sd_ind))
        esjd_err = np.hypot(out["esjd_se"][idx], _shared_stream_se(esjd_sd, n_seeds, esjd_sd_ind))
        print(f"[driver {target}] x={x}: acc {acc:.5f} vs {acc_ref} (3 s.e. {3 * acc_err:.5f}), esjd {esjd:.6f} vs {esjd_ref}")
        assert abs(acc - acc_ref) <= 3 * acc_err, (target, x, acc, acc_ref, acc_err)
        assert abs(esjd - esjd_ref) <= 3 * esjd_err + 0.02 * esjd_ref, (target, x, esjd, esjd_ref, esjd_err)


def test_pt_study_driver_matches_reference_data_points(tmp_path):
    """run_pt_study (the 30-rate sweep of experiment_pt_GPU.py:165-279: RoughCarpet +-15 d=20, 5e5 iterations, iterative
    ladder with the reference's high-precision settings, every ladder built by the native estimator) against the two
    recorded points of BASELINE.md section 2: actual swap rate 0.21961 / 0.28308 and PT-ESJD 0.008240 / 0.008344 (20 seeds,
    sd 0.0015 / 0.00012).  One run builds ONE ladder per rate, so it deviates from the 20-seed mean like a single seed
    does: 3 sd sqrt(1 + 1/20)."""
    _cuda()
    from rwm_pt_pytorch_b200.experiments import run_pt_study
    out = run_pt_study(20, "RoughCarpet", num_iters=500_000, swap_accept_max=0.5, seed=5, burn_in=1000, ladders_per_value=64,
                       out_dir=str(tmp_path))
    assert os.path.exists(out["filename"]) and len(out["acceptance_rates"]) == 30
    ref = {13: (0.229655, 0.21961, 0.00147, 0.008240, 0.000121), 17: (0.297241, 0.28308, 0.00146, 0.008344, 0.000107)}
    for idx, (rate, acc_ref, acc_sd, esjd_ref, esjd_sd) in ref.items():
        assert abs(out["swap_acceptance_rates_range"][idx] - rate) < 1e-5
        acc, esjd = out["acceptance_rates"][idx], out["expected_squared_jump_distances"][idx]
        print(f"[driver PT] target rate {rate}: swap {acc:.5f} vs {acc_ref}, pt_esjd {esjd:.6f} vs {esjd_ref}, K={out['ladder_sizes'][idx]}")
        assert abs(acc - acc_ref) <= 3 * acc_sd * np.sqrt(1 + 1 / 20), (idx, acc, acc_ref)
        assert abs(esjd - esjd_ref) <= 3 * esjd_sd * np.sqrt(1 + 1 / 20) + 0.02 * esjd_ref, (idx, esjd, esjd_ref)


def test_reference_average_seeds_tool_consumes_driver_output(tmp_path):
    """The reference's data/average_seeds.py, unmodified, on files written by run_rwm_study / run_pt_study (SURVEY 8f.2:
    "so that data/average_seeds.py and plot.py keep working")."""
    _cuda()
    tool = os.path.join(ROOT, "baseline", "_ref", "_reference_data_tools", "average_seeds.py")
    if not os.path.exists(tool):
        pytest.skip("baseline/_ref/_reference_data_tools missing (scripts/install_reference.py)")
    from rwm_pt_pytorch_b200.experiments import run_rwm_study, run_pt_study
    outs = [run_rwm_study(10, "EvenRosenbrock", num_iters=20000, var_max=0.6, seed=s, chains_per_value=4, out_dir=str(tmp_path))
            for s in (1, 2, 3)]
    pts = [run_pt_study(8, "ThreeMixture", num_iters=4000, seed=s, num_values=4, N_samples_swap_est=5000, iterative_tolerance=0.01,
                        ladders_per_value=4, swap_every=10, out_dir=str(tmp_path)) for s in (1, 2)]
    for pattern, runs, key in (("EvenRosenbrock_Normal_RWM_GPU_dim10_20000iters", outs, "var_value_range"),
                               ("ThreeMixture_PT_GPU_dim8_4000iters", pts, "swap_acceptance_rates_range")):
        p = subprocess.run([sys.executable, tool, "--pattern", pattern, "--data_dir", str(tmp_path), "--output_dir", str(tmp_path)],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
        made = [f for f in glob.glob(os.path.join(str(tmp_path), "*.json")) if "averaged" in os.path.basename(f).lower() and pattern in f]
        assert made, (os.listdir(str(tmp_path)), p.stdout[-800:])
        avg = json.load(open(made[0]))
        want = np.mean([r["acceptance_rates"] for r in runs], axis=0)
        np.testing.assert_allclose(avg["acceptance_rates"], want, rtol=1e-9)
        np.testing.assert_allclose(avg[key], runs[0][key], rtol=1e-12)
        assert avg["num_files_averaged"] == len(runs) and avg["dimension"] == runs[0]["dimension"]
        np.testing.assert_allclose(avg["max_esjd"], np.mean([r["max_esjd"] for r in runs]), rtol=1e-9)


def test_benchmark_performance_schema_and_rates():
    """MCMCSimulation_GPU.benchmark_performance (interfaces/simulation_gpu.py:252-311): same keys as the reference, one
    timing per sample size (and the reference's second "CPU" timing of the same object), positive rates, the simulation
    left reset-able with its original iteration count."""
    _cuda()
    from rwm_pt_pytorch_b200.interfaces import MCMCSimulation_GPU
    from rwm_pt_pytorch_b200.algorithms import RandomWalkMH_GPU_Optimized
    from rwm_pt_pytorch_b200.target_distributions import MultivariateNormalTorch
    sim = MCMCSimulation_GPU(dim=5, sigma=0.5, num_iterations=777, algorithm=RandomWalkMH_GPU_Optimized,
                             target_dist=MultivariateNormalTorch(5, device="cuda"), device="cuda", seed=1, burn_in=10)
    sizes = [500, 2000, 8000]
    r = sim.benchmark_performance(num_samples_list=sizes)
    assert set(r) == {"sample_sizes", "gpu_times", "gpu_samples_per_sec", "cpu_times", "cpu_samples_per_sec", "speedup"}
    assert list(r["sample_sizes"]) == sizes
    for k in ("gpu_times", "gpu_samples_per_sec", "cpu_times", "cpu_samples_per_sec", "speedup"):
        assert len(r[k]) == len(sizes) and all(v > 0 for v in r[k]), (k, r[k])
    assert all(abs(s / t - v) < 1e-6 * v for s, t, v in zip(sizes, r["gpu_times"], r["gpu_samples_per_sec"]))
    assert sim.num_iterations == 777
    r2 = sim.benchmark_performance(num_samples_list=[1000], compare_cpu=False)
    assert r2["cpu_times"] is None and r2["speedup"] is None and len(r2["gpu_times"]) == 1
    sim.reset()
    chain = sim.generate_samples(progress_bar=False)
    assert len(chain) == 777


def test_sharded_entry_points_equal_the_unsharded_run():
    """distributed.make_sharded / finish_sharded (the public multi-GPU entry: one process per GPU) on one process, and two
    hand-made shards with the ids make_sharded would assign: pooled statistics and gathered cold chains equal the single
    batch bit for bit (Philox subsequence = global chain id)."""
    _cuda()
    from rwm_pt_pytorch_b200 import distributed as D
    from rwm_pt_pytorch_b200.algorithms import ParallelTemperingRWM_GPU_Optimized as PT
    from rwm_pt_pytorch_b200.target_distributions import RoughCarpetDistributionTorch
    t = RoughCarpetDistributionTorch(20, device="cuda")
    kw = dict(geom_temp_spacing=True, swap_every=10, burn_in=50, store="cold", seed=9, swap_mode="reference")
    whole, shard = D.make_sharded(PT, 12, 20, 0.9, t, **kw)
    assert shard == (0, 12) and whole.num_ladders == 12
    whole.generate_samples(600)
    summary, samples = D.finish_sharded(whole, shard, 12, gather=True)
    parts = []
    for first, count in ((0, 7), (7, 5)):
        a = PT(20, 0.9, t, num_ladders=count, ladder_id_base=first, device="cuda", **kw)
        a.generate_samples(600)
        parts.append(a)
    both = torch.cat([p._batch.samples[:, :p._batch.rows_written()] for p in parts], dim=0)
    assert torch.equal(both, samples)
    acc = sum(int(p._batch.accept_count.sum()) for p in parts)
    assert summary["acceptance_rate"] == pytest.approx(acc / (12 * 8 * 600), rel=1e-12)
    sw = sum(int(p.num_swap_acceptances) for p in parts) / sum(int(p.num_swap_attempts) for p in parts)
    assert summary["swap_acceptance_rate"] == pytest.approx(sw, rel=1e-12)
