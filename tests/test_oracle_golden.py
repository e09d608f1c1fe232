"""CPU tests: the NumPy oracle (oracle/rwmpt_oracle.py) against every golden fixture produced by the
unmodified reference (tests/golden/make_golden.py).  These pin the oracle; the GPU parity tests then
compare the CUDA path with the oracle and with the same fixtures."""
import numpy as np
import pytest

from oracle import rwmpt_oracle as O
from tests._util import load_golden, golden_names

LOGP_RTOL = 2e-6   # torch-CPU and NumPy sum fp32 in different orders: last-ulp differences only


def _close_logp(a, b, rtol=LOGP_RTOL, atol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_inf | (np.abs(a - b) <= atol + rtol * np.abs(b))
    assert ok.all(), f"max abs diff {np.nanmax(np.abs(np.where(both_inf, 0, a - b)))}"


@pytest.mark.parametrize("name", golden_names("logp_"))
def test_log_density_known_answers(name):
    spec, g = load_golden(name)
    _close_logp(O.log_density(spec, g["x"]), g["logp"])
    for i in range(g["logp_single"].shape[0]):
        _close_logp(O.log_density(spec, g["x"][i]), g["logp_single"][i])


@pytest.mark.parametrize("tag", ["b1p0", "b0p25"])
def test_proposal_transforms(tag):
    _, g = load_golden(f"prop_normal_{tag}")
    assert O.normal_std(float(g["var"]), float(g["beta"])) == g["std"]
    np.testing.assert_array_equal(O.normal_increments(g["z"], float(g["var"]), float(g["beta"])), g["inc"])
    _, g = load_golden(f"prop_laplace_{tag}")
    np.testing.assert_array_equal(O.laplace_scale(g["var_vec"], float(g["beta"])), g["scale"])
    np.testing.assert_allclose(O.laplace_increments(g["u"], g["var_vec"], float(g["beta"])), g["inc"], rtol=3e-6, atol=1e-9)
    _, g = load_golden(f"prop_uniform_{tag}")
    assert O.uniform_radius(float(g["radius"]), float(g["beta"])) == g["eff_radius"]
    np.testing.assert_allclose(O.uniform_radius_increments(g["z"], g["u"], float(g["radius"]), float(g["beta"])),
                               g["inc"], rtol=3e-6, atol=1e-9)


@pytest.mark.parametrize("name", golden_names("rwm_"))
def test_rwm_matches_reference(name):
    spec, g = load_golden(name)
    T = g["increments"].shape[0]
    out = O.rwm_run(spec, g["x0"][None], g["beta"], g["increments"][:, None], g["uniforms"][:, None],
                    burn_in=int(g["burn_in"]))
    np.testing.assert_array_equal(out["decisions"][:, 0], g["decisions"])          # bit-exact decisions
    np.testing.assert_array_equal(out["chain"][:, 0], g["chain"])                  # states are x + inc: exact
    _close_logp(out["logp"][:, 0], g["logp"])
    assert int(out["accept_count"][0]) == int(g["num_acceptances"])
    assert out["acceptance_rate"][0] == pytest.approx(float(g["acceptance_rate"]), rel=1e-12)
    assert out["esjd"][0] == pytest.approx(float(g["esjd"]), rel=2e-5)
    assert O.esjd_from_chain(out["chain"][:, 0], int(g["burn_in"])) == pytest.approx(float(g["esjd"]), rel=1e-5)
    assert T == g["chain"].shape[0] - 1


@pytest.mark.parametrize("name", golden_names("pt_"))
def test_pt_matches_reference(name):
    spec, g = load_golden(name)
    betas = g["betas"]
    # the per-chain proposal std the reference bakes into its Cholesky factors (pt_rwm_gpu_optimized.py:453-455)
    std = np.sqrt((np.float32(1.0) * np.asarray([np.float32(float(g["var"]) / b) for b in betas], np.float32)))
    np.testing.assert_array_equal(std, g["chol_diag"])
    out = O.pt_run(spec, g["x0"][None], betas, g["increments"][:, None], g["uniforms"][:, None],
                   g["swap_uniforms"][:, None], int(g["swap_every"]), burn_in=int(g["burn_in"]), swap_mode="reference")
    np.testing.assert_array_equal(out["decisions"][:, 0], g["decisions"])
    np.testing.assert_array_equal(out["swap_decisions"][:, 0], g["swap_decisions"])
    np.testing.assert_array_equal(out["chain"][:, 0], g["states"])
    _close_logp(out["logp"][:, 0], g["logp"])
    assert int(out["swap_attempts"][0]) == int(g["num_swap_attempts"])
    assert int(out["swap_accepts"][0]) == int(g["num_swap_acceptances"])
    assert out["swap_acceptance_rate"][0] == pytest.approx(float(g["swap_acceptance_rate"]), rel=1e-12)
    assert out["pt_esjd"][0] == pytest.approx(float(g["pt_esjd"]), rel=1e-9)
    assert out["sq_beta_jump_sum"][0] == pytest.approx(float(g["squared_jump_distances"]), rel=1e-9)
    assert out["cold_esjd"][0] == pytest.approx(float(g["cold_esjd"]), rel=2e-5)


def test_pt_exchange_mode_conserves_states():
    """Textbook exchange permutes the ladder's states; the reference's copy k->j duplicates them."""
    spec, g = load_golden("pt_rough_carpet_pm4_d20_se3")
    K, d = g["x0"].shape
    rs = np.random.RandomState(3)
    x0 = rs.randn(1, K, d).astype(np.float32)
    T = 30
    zero_inc = np.zeros((T, 1, K, d), np.float32)       # chains never move: only the sweeps change the ladder
    args = (spec, x0, g["betas"], zero_inc, g["uniforms"][:T, None], g["swap_uniforms"][:T // 3, None], 3)
    ex = O.pt_run(*args, burn_in=0, swap_mode="exchange")
    ref = O.pt_run(*args, burn_in=0, swap_mode="reference")
    assert ex["swap_accepts"][0] > 0 and ref["swap_accepts"][0] > 0
    srt = lambda a: a[np.lexsort(a.T[::-1])]
    np.testing.assert_array_equal(srt(ex["final_state"][0]), srt(x0[0]))
    assert len(np.unique(ref["final_state"][0], axis=0)) < K
    # first sweep, pair 0: both modes take the same decision from the same pre-sweep state
    assert ex["swap_decisions"][0, 0, 0] == ref["swap_decisions"][0, 0, 0]


def test_geometric_ladder():
    assert O.geometric_ladder() == [1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 0.01]


def test_initial_state_rule():
    rs = np.random.RandomState(0)
    assert np.all(O.initial_state("RoughCarpetTorch", 5, rs) == 0)
    assert np.all(O.initial_state("ThreeMixtureTorchCustom", 5, rs) == 0)
    b = O.initial_state("IIDBetaTorch", 50, rs)
    assert b.dtype == np.float32 and (b > 0.2).all() and (b < 0.8).all()
    g = O.initial_state("IIDGammaTorch", 50, rs)
    assert abs(g.mean() - 5) < 0.01
    e = O.initial_state("EvenRosenbrockTorch", 50, rs)
    assert np.abs(e).max() < 1e-7


def test_numpy_cpu_sampler_restatement():
    """algorithms/rwm.py restated: same NumPy RNG call order => same chain as the reference run."""
    _, g = load_golden("numpy_rwm_c1")
    out = O.numpy_rwm_cpu(O.rough_carpet_density_cpu, 20, float(g["var"]), int(g["n_steps"]), seed=1)
    np.testing.assert_allclose(out["chain"][:4], g["chain_head"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(out["chain"][-2:], g["chain_tail"], rtol=1e-9, atol=1e-12)
    assert out["acceptance_rate"] == pytest.approx(float(g["acceptance_rate"]), rel=1e-12)
    assert out["esjd"] == pytest.approx(float(g["esjd"]), rel=1e-9)


@pytest.mark.parametrize("name", golden_names("ladder_"))
def test_swap_probability_estimator_and_ladder_match_reference(name):
    """The oracle's restatement of the iterative-ladder estimator (pt_rwm_gpu_optimized.py:356-368 + the targets'
    draw_samples_torch) against the reference's own estimates on a grid of (beta, beta*) pairs, and the ladder the
    restated recursion (:283-426) builds against the reference's ladder, rung by rung."""
    spec, g = load_golden(name)
    d = int(g["dim"])
    rs = np.random.RandomState(7)
    N = 40_000
    for bc, bs, ref, se in zip(g["pair_beta"], g["pair_beta_star"], g["pair_estimate"], g["pair_se"]):
        est, se_o = O.swap_probability(spec, d, float(bc), float(bs), N, rs)
        assert abs(est - ref) <= 4.5 * np.hypot(se, se_o) + 1e-4, (bc, bs, est, ref, se, se_o)
    # every rung the reference accepted had an estimate within `tolerance` of the target (it used n_est samples): with the
    # oracle's estimator at the reference's rungs the rate must be the target within tolerance + Monte-Carlo error
    lad = g["ladder"]
    k_pairs = int(g["n_ladder_pairs"])
    for k in range(k_pairs - 1):            # the last rung is the forced beta_min
        est, se_o = O.swap_probability(spec, d, float(lad[k]), float(lad[k + 1]), N, rs)
        se_ref = np.sqrt(0.25 / float(g["n_est"]))      # the reference's own estimate: n_est samples, variance <= 1/4
        assert abs(est - float(g["target_rate"])) <= float(g["tolerance"]) + 4 * np.hypot(se_o, se_ref), (k, est)
    # the restated recursion rebuilds the ladder: same number of rungs (+-1), rungs within 12 % of the reference's
    mine = O.iterative_ladder(lambda b, bs, n: O.swap_probability(spec, d, b, bs, n, rs)[0], target_rate=float(g["target_rate"]),
                              n_samples=int(g["n_est"]))
    assert abs(len(mine) - len(lad)) <= 1, (mine, lad.tolist())
    m = min(len(mine), len(lad)) - 1
    np.testing.assert_allclose(mine[:m], lad[:m], rtol=0.12)
