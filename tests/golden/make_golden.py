#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference (`aidanmrli/rwm-pt-pytorch`) is imported as-is with a stub `matplotlib` module
(it imports pyplot at module scope but never uses it on this path) and run on the torch CPU
device, so TF32 never perturbs `torch.bmm` / `torch.matmul`.  For every case we
capture the exact randomness the reference consumed and what it did with it:

  rwm_*.npz : x0, increments (T,d), uniforms (T,), decisions (T,), chain (T+1,d), logp (T+1,),
              acceptance_rate, esjd, plus the target's parameters (keys `spec_*`).
  pt_*.npz  : x0, betas, post-bmm increments (T,K,d), uniforms (T,K), swap_uniforms (R,K-1),
              decisions (T,K), swap_decisions (R,K-1), states (T+1,K,d), logp (T+1,K), swap stats.
  logp_*.npz: random points and the reference's `log_density` there.
  prop_*.npz: raw torch.rand/randn draws and the proposal plugin's `sample()` output.
  numpy_rwm_c1.npz : the NumPy CPU sampler (algorithms/rwm.py) on BASELINE config 1, shortened.

Nothing here is imported by the product or by the GPU-side tests; the fixtures are.
"""
import io
import os
import sys
import types
import contextlib

import numpy as np

REF = os.environ.get("RWMPT_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

# --- import the reference unmodified -------------------------------------------------------
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)

import torch  # noqa: E402

import algorithms.rwm_gpu_optimized as ref_rwm_mod  # noqa: E402
import algorithms.pt_rwm_gpu_optimized as ref_pt_mod  # noqa: E402
from algorithms.rwm import RandomWalkMH  # noqa: E402
from proposal_distributions import NormalProposal, LaplaceProposal, UniformRadiusProposal  # noqa: E402
import target_distributions as td  # noqa: E402

CPU = torch.device("cpu")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def t2n(t):
    return t.detach().cpu().numpy()


# --- reference target object -> oracle spec -------------------------------------------------
def spec_of(t) -> dict:
    n = type(t).__name__
    if n == "RoughCarpetDistributionTorch":
        s = dict(family="rough_carpet", modes=t2n(t.modes), log_weights=t2n(t.log_weights),
                 log_sqrt_2pi=t2n(t.log_sqrt_2pi))
        if hasattr(t, "scaling_factors"):
            s["scaling"] = t2n(t.scaling_factors)
        return s
    if n == "ThreeMixtureDistributionTorch":
        s = dict(family="three_mixture", means=t2n(t.means), log_weights=t2n(t.log_mixing_weights))
        if t.scaling_arg_from_constructor:
            s["scaling"] = t2n(t.scaling_factors)
            s["log_jacobian"] = t2n(t.log_jacobian)
            s["c1"] = np.full(3, t2n(t.base_log_norm_const_for_scaled), dtype=np.float32)
        else:
            s["c1"] = t2n(t.log_norm_consts)
        return s
    if n in ("FullRosenbrockTorch", "EvenRosenbrockTorch"):
        fam = "full_rosenbrock" if n.startswith("Full") else "even_rosenbrock"
        return dict(family=fam, a=t2n(t.a_coeff), b=t2n(t.b_coeff), mu=t2n(t.mu))
    if n == "HybridRosenbrockTorch":
        return dict(family="hybrid_rosenbrock", a=t2n(t.a_coeff), b=t2n(t.b_coeff), mu=t2n(t.mu),
                    n1=np.int64(t.n1), n2=np.int64(t.n2))
    if n == "NealFunnelTorch":
        return dict(family="neal_funnel", mu_v=t2n(t.mu_v), sigma_v_sq=t2n(t.sigma_v_sq), mu_z=t2n(t.mu_z),
                    log_sigma_v_sq=t2n(t.log_sigma_v_sq), log_2pi=t2n(t.log_2_pi), dm1=t2n(t.D_minus_1_tensor))
    if n == "HypercubeTorch":
        return dict(family="hypercube", left=t2n(t.left_boundary), right=t2n(t.right_boundary),
                    log_uniform_density=t2n(t.log_uniform_density))
    if n == "IIDGammaTorch":
        return dict(family="iid_gamma", shape=t2n(t.shape), scale=t2n(t.scale), log_norm_const=t2n(t.log_norm_const))
    if n == "IIDBetaTorch":
        return dict(family="iid_beta", alpha=t2n(t.alpha), beta=t2n(t.beta), log_norm_const=t2n(t.log_norm_const))
    if n == "ScaledMultivariateNormalTorch":
        return dict(family="scaled_mvn", c=t2n(t.scaling_factors), log_norm_const=t2n(t.log_norm_const))
    if n == "MultivariateNormalTorch":
        if not torch.allclose(torch.diag(torch.diagonal(t.cov_inv)), t.cov_inv):
            return dict(family="mvn_dense", mean=t2n(t.mean), cov_inv=t2n(t.cov_inv), log_norm_const=t2n(t.log_norm_const))
        prec = t2n(torch.diagonal(t.cov_inv))
        return dict(family="mvn_diag", mean=t2n(t.mean), prec=prec, log_norm_const=t2n(t.log_norm_const))
    if n == "SuperFunnelTorch":
        rec = torch.cat([torch.cat([torch.full((t.Y_data[j].shape[0], 1), float(j)), t.Y_data[j].reshape(-1, 1), t.X_data[j]], dim=1)
                         for j in range(t.J)], dim=0)
        return dict(family="super_funnel", J=np.int64(t.J), K=np.int64(t.K), records=t2n(rec), hyper_var=t2n(t.prior_hypermean_var),
                    log_hyper_var=t2n(t.log_prior_hypermean_var), tau_scale=t2n(t.prior_tau_scale),
                    log_tau_scale=t2n(t.log_prior_tau_scale), log_2pi=t2n(t.log_2_pi), log_2=t2n(t.log_2), log_pi=t2n(t.log_pi))
    raise ValueError(n)


def save(name, spec, **arrays):
    out = {f"spec_{k}": np.asarray(v) for k, v in spec.items()}
    out.update({k: np.asarray(v) for k, v in arrays.items()})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {name}.npz ({os.path.getsize(path) / 1024:.1f} KiB)")


# --- targets used by the cases --------------------------------------------------------------
DENSE_MEAN = [0.3 * ((-1) ** i) * (i % 4) for i in range(12)]
DENSE_COV = [[(0.7 ** abs(i - j)) * (1.0 + 0.1 * min(i, j)) for j in range(12)] for i in range(12)]


def make_super_funnel(cls, J=5, K=3, n_per=20):
    g = torch.Generator().manual_seed(42)
    X, Y = [], []
    for _ in range(J):
        Xj = torch.randn(n_per, K, generator=g)
        X.append(Xj)
        Y.append(torch.bernoulli(torch.sigmoid(0.5 * torch.sum(Xj, dim=1)), generator=g))
    return cls(J, K, X, Y, prior_hypermean_std=10.0, prior_tau_scale=2.5, device=CPU)


def make_targets():
    torch.manual_seed(1234)  # fixes the random scaling factors of the *Scaled targets
    d = {}
    d["rough_carpet_d20"] = td.RoughCarpetDistributionTorch(20, device=CPU)
    d["rough_carpet_pm4_d20"] = td.RoughCarpetDistributionTorch(20, device=CPU, mode_centers=[-4.0, 0.0, 4.0])
    d["rough_carpet_scaled_d6"] = td.RoughCarpetDistributionTorch(6, scaling=True, device=CPU)
    d["three_mixture_d10"] = td.ThreeMixtureDistributionTorch(10, device=CPU)
    d["three_mixture_pm15_d50"] = td.ThreeMixtureDistributionTorch(
        50, device=CPU, mode_centers=[[-15.0] + [0.0] * 49, [0.0] * 50, [15.0] + [0.0] * 49])
    d["three_mixture_scaled_d7"] = td.ThreeMixtureDistributionTorch(7, scaling=True, device=CPU)
    d["full_rosenbrock_d20"] = td.FullRosenbrockTorch(20, device=CPU)
    d["full_rosenbrock_d3"] = td.FullRosenbrockTorch(3, device=CPU)
    d["even_rosenbrock_d10"] = td.EvenRosenbrockTorch(10, device=CPU)
    d["even_rosenbrock_d20"] = td.EvenRosenbrockTorch(20, device=CPU)
    d["even_rosenbrock_d30"] = td.EvenRosenbrockTorch(30, device=CPU)
    d["hybrid_rosenbrock_n3x5"] = td.HybridRosenbrockTorch(3, 5, device=CPU)
    d["hybrid_rosenbrock_n4x2"] = td.HybridRosenbrockTorch(4, 2, device=CPU)
    d["neal_funnel_d10"] = td.NealFunnelTorch(10, device=CPU)
    d["neal_funnel_d1"] = td.NealFunnelTorch(1, device=CPU)
    d["hypercube_pm1_d5"] = td.HypercubeTorch(5, left_boundary=-1, right_boundary=1, device=CPU)
    d["hypercube_01_d4"] = td.HypercubeTorch(4, device=CPU)
    d["iid_gamma_d8"] = td.IIDGammaTorch(8, shape=2, scale=3, device=CPU)
    d["iid_beta_d8"] = td.IIDBetaTorch(8, alpha=2, beta=3, device=CPU)
    d["scaled_mvn_d12"] = td.ScaledMultivariateNormalTorch(12, device=CPU)
    d["mvn_identity_d50"] = td.MultivariateNormalTorch(50, device=CPU)
    d["mvn_diag_d6"] = td.MultivariateNormalTorch(
        6, mean=[0.5, -1.0, 0.0, 2.0, 0.25, -0.75], cov=np.diag([0.5, 2.0, 1.0, 4.0, 0.25, 1.5]).tolist(), device=CPU)
    # BASELINE config 5's shape (13 coordinates x 8 lanes with padding on the CUDA side); unscaled targets draw no randomness,
    # so appending them leaves every earlier fixture bit-identical
    d["full_rosenbrock_d100"] = td.FullRosenbrockTorch(100, device=CPU)
    d["neal_funnel_d100"] = td.NealFunnelTorch(100, device=CPU)
    # dense-covariance Gaussian (AR(1)-like covariance, not diagonal) and the hierarchical logistic
    # "super funnel" on the synthetic data of experiment_RWM_GPU.py:95-120 (own generator: the global stream is untouched)
    d["mvn_dense_d12"] = td.MultivariateNormalTorch(12, mean=DENSE_MEAN, cov=DENSE_COV, device=CPU)
    d["super_funnel_j5k3"] = make_super_funnel(td.SuperFunnelTorch)
    return d


# --- log-density known answers ----------------------------------------------------------------
def gen_logp(name, target):
    g = torch.Generator().manual_seed(99)
    dim = target.dim
    n = 96
    tn = type(target).__name__
    if tn == "IIDBetaTorch":
        x = torch.rand(n, dim, generator=g) * 1.2 - 0.1            # some rows leave (0,1)
    elif tn == "IIDGammaTorch":
        x = torch.rand(n, dim, generator=g) * 12.0 - 0.5
    elif tn == "HypercubeTorch":
        x = torch.rand(n, dim, generator=g) * 2.6 - 1.3
        x[0] = target.left_boundary; x[1] = target.right_boundary     # inclusive boundaries
    elif "Rosenbrock" in tn:
        x = torch.randn(n, dim, generator=g) * 1.5 + 0.5
    elif tn == "NealFunnelTorch":
        x = torch.randn(n, dim, generator=g) * 3.0
    elif tn == "SuperFunnelTorch":
        x = torch.randn(n, dim, generator=g) * 1.5
        x[:, -2:] = x[:, -2:].abs() + 0.05            # taus in the support ...
        x[:6, -1] = torch.tensor([-0.3, 0.0, 1e-10, 2.0, 0.5, 1.0])   # ... except a few rows (-> -inf)
        x[3:6, -2] = torch.tensor([-1.0, 0.0, 3.0])
    else:
        x = torch.randn(n, dim, generator=g) * 6.0
    x = x.to(torch.float32)
    lp_batch = target.log_density(x)
    lp_single = torch.stack([target.log_density(x[i]).reshape(()) for i in range(8)])
    save(f"logp_{name}", spec_of(target), x=t2n(x), logp=t2n(lp_batch), logp_single=t2n(lp_single), dim=dim)


# --- proposal plugin known answers -------------------------------------------------------------
def gen_proposals():
    n, d = 257, 9
    for beta in (1.0, 0.25):
        tag = f"b{str(beta).replace('.', 'p')}"
        p = NormalProposal(d, 0.37, beta, CPU, torch.float32)
        torch.manual_seed(5); inc = p.sample(n)
        torch.manual_seed(5); z = torch.randn((n, d))
        save(f"prop_normal_{tag}", {}, z=t2n(z), inc=t2n(inc), var=0.37, beta=beta, std=t2n(p.std_dev))
        var_vec = torch.linspace(0.05, 1.3, d)
        p = LaplaceProposal(d, var_vec, beta, CPU, torch.float32)
        torch.manual_seed(6); inc = p.sample(n)
        torch.manual_seed(6); u = torch.rand((n, d))
        save(f"prop_laplace_{tag}", {}, u=t2n(u), inc=t2n(inc), var_vec=t2n(var_vec), beta=beta,
             scale=t2n(p.scale_vector))
        p = UniformRadiusProposal(d, 1.7, beta, CPU, torch.float32)
        torch.manual_seed(7); inc = p.sample(n)
        torch.manual_seed(7); z = torch.randn((n, d)); u = torch.rand((n, 1))
        save(f"prop_uniform_{tag}", {}, z=t2n(z), u=t2n(u), inc=t2n(inc), radius=1.7, beta=beta,
             eff_radius=t2n(p.effective_radius))


# --- RWM: run RandomWalkMH_GPU_Optimized on the CPU device and capture everything ---------------
def gen_rwm(name, target, n_samples, burn_in, seed, var=None, proposal=None, beta=1.0, x0=None):
    np.random.seed(seed)  # initial state of the "else" branch comes from NumPy's global RNG
    with quiet():
        algo = ref_rwm_mod.RandomWalkMH_GPU_Optimized(
            dim=target.dim, var=var, target_dist=target, beta=beta, burn_in=burn_in, device="cpu",
            pre_allocate_steps=n_samples, proposal_distribution=proposal)
    if x0 is not None:       # generate_samples starts from chain[-1] (rwm_gpu_optimized.py:431-434)
        algo.chain = [np.asarray(x0, dtype=np.float64)]
    cap = {"acc": []}
    orig_pre = algo._precompute_all_randoms

    def pre(total):
        orig_pre(total)
        cap["inc"] = algo.precomputed_increments.clone()
        cap["u"] = algo.precomputed_random_vals.clone()

    algo._precompute_all_randoms = pre
    orig_fused = ref_rwm_mod.ultra_fused_mcmc_step_basic

    def fused(*a):
        out = orig_fused(*a)
        cap["acc"].append(bool(out[2].item()))
        return out

    ref_rwm_mod.ultra_fused_mcmc_step_basic = fused
    try:
        torch.manual_seed(seed)
        with quiet():
            samples = algo.generate_samples(n_samples)
    finally:
        ref_rwm_mod.ultra_fused_mcmc_step_basic = orig_fused
    chain = algo.get_chain_gpu()
    logp = algo.get_log_densities_gpu()
    assert chain.shape[0] == burn_in + n_samples + 1 and samples.shape[0] == n_samples
    save(name, spec_of(target),
         x0=np.asarray(algo.chain[0], dtype=np.float32), beta=np.float32(beta), burn_in=burn_in,
         increments=t2n(cap["inc"]), uniforms=t2n(cap["u"]), decisions=np.asarray(cap["acc"], dtype=np.uint8),
         chain=t2n(chain), logp=t2n(logp), acceptance_rate=algo.acceptance_rate,
         num_acceptances=algo.num_acceptances, esjd=algo.expected_squared_jump_distance_gpu(),
         algo_name=algo.get_name(), target_name=target.get_name(), dim=target.dim)


# --- PT: run ParallelTemperingRWM_GPU_Optimized on the CPU device and capture everything -------
def gen_pt(name, target, n_samples, burn_in, seed, var, swap_every, beta_ladder=None):
    np.random.seed(seed)
    with quiet():
        algo = ref_pt_mod.ParallelTemperingRWM_GPU_Optimized(
            dim=target.dim, var=var, target_dist=target, beta_ladder=beta_ladder,
            geom_temp_spacing=beta_ladder is None, swap_every=swap_every, burn_in=burn_in, device="cpu",
            pre_allocate_steps=n_samples)
    K = algo.num_chains
    cap = {"inc": [], "acc": [], "states": [algo.current_states.clone()], "logp": [algo.current_log_densities.clone()],
           "swaps": {}}
    x0 = algo.current_states.clone()
    orig_bmm = ref_pt_mod.batch_matrix_multiply_increments
    orig_fused = ref_pt_mod.ultra_fused_parallel_mcmc_step
    orig_exec = ref_pt_mod.fused_swap_execution_no_clone

    def bmm(chol, raw):
        out = orig_bmm(chol, raw)
        cap["inc"].append(out.clone())
        return out

    def fused(*a):
        out = orig_fused(*a)
        cap["acc"].append(out[2].clone())
        return out

    def swap_exec(states, lps, j, k):
        cap["swaps"].setdefault(algo.step_counter, []).append(j)
        return orig_exec(states, lps, j, k)

    orig_step = algo.step

    def step(step_index=None):
        if "u" not in cap:
            cap["u"] = algo.precomputed_mcmc_randoms.clone()
            cap["su"] = algo.precomputed_swap_randoms.clone()
        orig_step(step_index=step_index)
        cap["states"].append(algo.current_states.clone())
        cap["logp"].append(algo.current_log_densities.clone())

    algo.step = step
    ref_pt_mod.batch_matrix_multiply_increments = bmm
    ref_pt_mod.ultra_fused_parallel_mcmc_step = fused
    ref_pt_mod.fused_swap_execution_no_clone = swap_exec
    try:
        torch.manual_seed(seed)
        with quiet():
            cold = algo.generate_samples(n_samples)
    finally:
        ref_pt_mod.batch_matrix_multiply_increments = orig_bmm
        ref_pt_mod.ultra_fused_parallel_mcmc_step = orig_fused
        ref_pt_mod.fused_swap_execution_no_clone = orig_exec
    T = burn_in + n_samples
    rounds = [s for s in range(1, T + 1) if s % swap_every == 0 and s > burn_in]
    R = len(rounds)
    swap_dec = np.zeros((R, K - 1), dtype=np.uint8)
    for r, s in enumerate(rounds):
        for j in cap["swaps"].get(s, []):
            swap_dec[r, j] = 1
    su = t2n(cap["su"])[: R * (K - 1)].reshape(R, K - 1)
    assert algo.num_swap_attempts == R * (K - 1)
    assert cold.shape[0] == n_samples
    save(name, spec_of(target),
         x0=t2n(x0), betas=np.asarray(algo.beta_ladder, dtype=np.float64), burn_in=burn_in, swap_every=swap_every,
         var=var, increments=t2n(torch.stack(cap["inc"])), uniforms=t2n(cap["u"])[:T], swap_uniforms=su,
         decisions=t2n(torch.stack(cap["acc"])).astype(np.uint8), swap_decisions=swap_dec,
         states=t2n(torch.stack(cap["states"])), logp=t2n(torch.stack(cap["logp"])),
         chol_diag=t2n(torch.diagonal(algo.proposal_covs_chol, dim1=1, dim2=2)[:, 0]),
         swap_acceptance_rate=algo.swap_acceptance_rate, num_swap_attempts=algo.num_swap_attempts,
         num_swap_acceptances=algo.num_swap_acceptances, pt_esjd=algo.pt_esjd,
         squared_jump_distances=algo.squared_jump_distances,
         cold_esjd=algo.expected_squared_jump_distance_gpu(), algo_name=algo.get_name(),
         target_name=target.get_name(), dim=target.dim)


# --- NumPy CPU sampler, BASELINE config 1 (shortened) --------------------------------------------
def gen_numpy_c1(n_steps=4000):
    target = td.RoughCarpetDistribution(20)
    algo = RandomWalkMH(20, 2.38 ** 2 / 20, target)
    np.random.seed(1)  # after construction, as interfaces/simulation.py:26-33 does
    for _ in range(n_steps):
        algo.step()
    chain = np.asarray(algo.chain)
    save("numpy_rwm_c1", {}, n_steps=n_steps, var=2.38 ** 2 / 20, acceptance_rate=algo.acceptance_rate,
         chain_head=chain[:4], chain_tail=chain[-2:],
         esjd=float(np.mean(np.sum((chain[1:] - chain[:-1]) ** 2, axis=1))))


# --- iterative temperature ladder: the reference's own ladder and its swap-probability estimator --------
def gen_ladder(name, target, target_rate, n_est, seed):
    """ladder_*.npz: (i) the ladder `_construct_iterative_ladder` (pt_rwm_gpu_optimized.py:283-426) builds on the torch CPU
    device, with the estimate it accepted for every rung; (ii) the reference's estimator itself -- draw_samples_torch at both
    temperatures, log_density, mean(exp(clamp_max(.,0))) (:356-368) -- on a grid of (beta, beta*) pairs with 4e5 samples
    each, with its Monte-Carlo standard error.  The CUDA estimator is checked against (ii) pair by pair and must rebuild
    (i) rung by rung."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        algo = ref_pt_mod.ParallelTemperingRWM_GPU_Optimized(
            dim=target.dim, var=2.38 ** 2 / target.dim, target_dist=target, iterative_temp_spacing=True,
            swap_acceptance_rate=target_rate, N_samples_swap_est=n_est, swap_every=10, device="cpu", pre_allocate_steps=10)
    ladder = np.asarray(algo.beta_ladder, dtype=np.float64)
    betas, stars, est, se = [], [], [], []
    g = torch.Generator().manual_seed(seed + 1)
    pairs = [(float(ladder[k]), float(ladder[k + 1])) for k in range(len(ladder) - 1)]
    for b in (1.0, 0.3, 0.05):                       # off-ladder pairs: ratios from near-certain to rare swaps
        for ratio in (0.95, 0.8, 0.5):
            pairs.append((b, b * ratio))
    N = 400_000
    for bc, bs in pairs:
        torch.manual_seed(int(torch.randint(0, 2 ** 31, (1,), generator=g)))
        xs = target.draw_samples_torch(N, bs)
        xc = target.draw_samples_torch(N, bc)
        log_r = (bc - bs) * (target.log_density(xs) - target.log_density(xc))
        p = torch.exp(torch.clamp_max(log_r, 0.0)).double()
        betas.append(bc); stars.append(bs); est.append(float(p.mean())); se.append(float(p.std() / np.sqrt(N)))
    save(name, spec_of(target), ladder=ladder, target_rate=target_rate, n_est=n_est, tolerance=0.005,
         pair_beta=np.asarray(betas), pair_beta_star=np.asarray(stars), pair_estimate=np.asarray(est), pair_se=np.asarray(se),
         n_ladder_pairs=len(ladder) - 1, dim=target.dim)
    print(f"  {name}: {len(ladder)} rungs", np.round(ladder, 5).tolist())


def main():
    T = make_targets()
    for name, t in T.items():
        gen_logp(name, t)
    gen_proposals()

    d = 20
    gen_rwm("rwm_rough_carpet_d20_normal", T["rough_carpet_d20"], 700, 100, seed=1, var=2.38 ** 2 / d)
    gen_rwm("rwm_rough_carpet_pm4_d20_beta0p5", T["rough_carpet_pm4_d20"], 400, 0, seed=2, var=1.2, beta=0.5)
    gen_rwm("rwm_rough_carpet_scaled_d6", T["rough_carpet_scaled_d6"], 400, 50, seed=3, var=0.8)
    gen_rwm("rwm_three_mixture_d10", T["three_mixture_d10"], 500, 50, seed=4, var=2.38 ** 2 / 10)
    gen_rwm("rwm_three_mixture_scaled_d7", T["three_mixture_scaled_d7"], 400, 0, seed=5, var=0.5)
    gen_rwm("rwm_even_rosenbrock_d10", T["even_rosenbrock_d10"], 600, 100, seed=6, var=0.297436 ** 2 / 10)
    gen_rwm("rwm_even_rosenbrock_d20", T["even_rosenbrock_d20"], 600, 100, seed=7, var=0.297436 ** 2 / 20)
    gen_rwm("rwm_even_rosenbrock_d30", T["even_rosenbrock_d30"], 500, 100, seed=8, var=0.297436 ** 2 / 30)
    gen_rwm("rwm_full_rosenbrock_d20", T["full_rosenbrock_d20"], 600, 100, seed=9, var=0.340769 ** 2 / 20)
    gen_rwm("rwm_full_rosenbrock_d3", T["full_rosenbrock_d3"], 300, 0, seed=10, var=0.1)
    gen_rwm("rwm_hybrid_rosenbrock_n3x5", T["hybrid_rosenbrock_n3x5"], 400, 50, seed=11, var=0.02)
    gen_rwm("rwm_hybrid_rosenbrock_n4x2", T["hybrid_rosenbrock_n4x2"], 300, 0, seed=12, var=0.02)
    gen_rwm("rwm_neal_funnel_d10", T["neal_funnel_d10"], 600, 100, seed=13, var=1.699744 ** 2 / 10)
    gen_rwm("rwm_neal_funnel_d1", T["neal_funnel_d1"], 200, 0, seed=14, var=2.0)
    gen_rwm("rwm_hypercube_pm1_d5", T["hypercube_pm1_d5"], 400, 0, seed=15, var=0.05)
    gen_rwm("rwm_hypercube_01_d4", T["hypercube_01_d4"], 300, 0, seed=16, var=0.05)   # starts outside the support
    gen_rwm("rwm_iid_gamma_d8", T["iid_gamma_d8"], 400, 50, seed=17, var=2.0)
    gen_rwm("rwm_iid_beta_d8", T["iid_beta_d8"], 400, 50, seed=18, var=0.02)
    gen_rwm("rwm_scaled_mvn_d12", T["scaled_mvn_d12"], 400, 50, seed=19, var=0.6)
    gen_rwm("rwm_mvn_diag_d6", T["mvn_diag_d6"], 300, 0, seed=20, var=0.9)
    # the three proposal plugins through the `proposal_distribution=` path (MCMCSimulation_GPU's route)
    gen_rwm("rwm_mvn_identity_d50_laplace", T["mvn_identity_d50"], 400, 50, seed=21,
            proposal=LaplaceProposal(50, torch.full((50,), 2.38 ** 2 / 50), 1.0, CPU, torch.float32))
    gen_rwm("rwm_mvn_identity_d50_uniform", T["mvn_identity_d50"], 400, 50, seed=22,
            proposal=UniformRadiusProposal(50, 1.0, 1.0, CPU, torch.float32))
    gen_rwm("rwm_three_mixture_pm15_d50_laplace_beta0p25", T["three_mixture_pm15_d50"], 300, 0, seed=23, beta=0.25,
            proposal=LaplaceProposal(50, torch.full((50,), 2.38 ** 2 / 50), 0.25, CPU, torch.float32))
    gen_rwm("rwm_rough_carpet_d20_uniform", T["rough_carpet_d20"], 400, 50, seed=24,
            proposal=UniformRadiusProposal(20, 2.5, 1.0, CPU, torch.float32))

    gen_rwm("rwm_mvn_dense_d12", T["mvn_dense_d12"], 400, 50, seed=27, var=0.5)
    sf0 = np.zeros(T["super_funnel_j5k3"].dim); sf0[-2:] = 1.0     # the default start (1e-8 N(0,1)) sits in the funnel's neck
    gen_rwm("rwm_super_funnel_j5k3", T["super_funnel_j5k3"], 500, 0, seed=28, var=0.02, x0=sf0)
    # BASELINE config 5: d = 100, at one point of each variance sweep (experiment_RWM_GPU.py:202-218: variance = x^2 / dim)
    gen_rwm("rwm_full_rosenbrock_d100", T["full_rosenbrock_d100"], 500, 100, seed=25, var=0.340769 ** 2 / 100)
    gen_rwm("rwm_neal_funnel_d100", T["neal_funnel_d100"], 500, 100, seed=26, var=1.699744 ** 2 / 100)

    gen_pt("pt_rough_carpet_d20_geom", T["rough_carpet_d20"], 1000, 200, seed=1, var=0.9, swap_every=10)
    gen_pt("pt_rough_carpet_pm4_d20_se3", T["rough_carpet_pm4_d20"], 300, 0, seed=2, var=2.38 ** 2 / 20, swap_every=3)
    gen_pt("pt_three_mixture_d10_k5", T["three_mixture_d10"], 600, 100, seed=3, var=2.38 ** 2 / 10, swap_every=5,
           beta_ladder=[1.0, 0.6, 0.3, 0.1, 0.02])
    gen_pt("pt_even_rosenbrock_d10_k3", T["even_rosenbrock_d10"], 400, 40, seed=4, var=0.01, swap_every=7,
           beta_ladder=[1.0, 0.5, 0.2])
    gen_pt("pt_neal_funnel_d10_k13", T["neal_funnel_d10"], 300, 30, seed=5, var=0.5, swap_every=4,
           beta_ladder=[float(b) for b in np.geomspace(1.0, 0.01, 13)])
    gen_pt("pt_iid_gamma_d8_k2", T["iid_gamma_d8"], 300, 0, seed=6, var=1.0, swap_every=2, beta_ladder=[1.0, 0.3])
    gen_numpy_c1()
    # ladder_*: built last so that the fixtures above keep their RNG streams
    gen_ladder("ladder_rough_carpet_pm15_d20", td.RoughCarpetDistributionTorch(20, device=CPU, mode_centers=[-15.0, 0.0, 15.0]),
               0.234, 20000, seed=31)
    gen_ladder("ladder_three_mixture_pm15_d30", td.ThreeMixtureDistributionTorch(
        30, device=CPU, mode_centers=[[-15.0] + [0.0] * 29, [0.0] * 30, [15.0] + [0.0] * 29], mode_weights=[1 / 3, 1 / 3, 1 / 3]),
        0.234, 20000, seed=32)


if __name__ == "__main__":
    main()
