"""CPU oracle for the RWM / PT-RWM sampling hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy (float32) restatement of the reference's algorithm
for the path named in BASELINE.json (`algorithms/rwm_gpu_optimized.py`,
`algorithms/pt_rwm_gpu_optimized.py`, `proposal_distributions/*`, the
`log_density` methods of `target_distributions/*_torch.py` and the initial
state rule of `interfaces/metropolis.py`).  Every function cites the reference
file:line it follows (paths relative to the reference repository root).

It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing
under `rwm_pt_pytorch_b200/` imports it and the product fails loudly when the
CUDA library is missing.

Pinning.  The reference's own tests hold no golden vectors for this path
(SURVEY.md section 4), so the pins are outputs of the *unmodified reference run in the
authoring container* (torch CPU device, so no TF32): `tests/golden/make_golden.py`
imports `/root/reference`, captures the exact randomness it consumed and its
per-step decisions / states, and writes `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks this restatement against every one of
those fixtures (decisions bit-exact, states exact, log-densities to 2e-6
relative: torch's and NumPy's fp32 reduction orders differ in the last ulp).
BASELINE config 4 (PT with Laplace / UniformRadius proposals) has no reference
implementation (`pt_rwm_gpu_optimized.py:114` takes `var` only); for it the
oracle composes the reference's proposal transforms with the reference's PT
step, and that composition is "parity unpinned" by the reference.

Conventions.  Everything is vectorised over independent chains (RWM) or
independent ladders (PT) -- the reference runs one chain / one ladder per Python
object, the product runs thousands at once -- and loops over MCMC steps in
Python exactly like the reference does.  All arithmetic is float32, in the
operation order listed in SURVEY.md section 8(a').
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np

F32 = np.float32
_NEG_INF = F32(-np.inf)


def _f(x):
    return np.asarray(x, dtype=np.float32)


def _logsumexp_last(t: np.ndarray) -> np.ndarray:
    """torch.logsumexp(t, dim=-1): m + log(sum(exp(t - m))), m = max (0 where |m| = inf)."""
    m = np.max(t, axis=-1, keepdims=True)
    m_safe = np.where(np.isinf(m), F32(0.0), m).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s = np.sum(np.exp(t - m_safe), axis=-1, dtype=np.float32)
        return (np.log(s) + m_safe[..., 0]).astype(np.float32)


# --------------------------------------------------------------------------------------
# Target log-densities (batched: x is (B, d) float32 -> (B,) float32)
# --------------------------------------------------------------------------------------

def logp_rough_carpet(x, spec):
    """`RoughCarpetDistributionTorch.log_density`, target_distributions/multimodal_torch.py:470-510."""
    x = _f(x)
    scaling = spec.get("scaling")
    if scaling is not None:
        xs = x * _f(scaling)                                  # :483-485
        log_jac = np.sum(np.log(_f(scaling)), dtype=np.float32)  # :486 (recomputed per call)
    else:
        xs = x
        log_jac = F32(0.0)
    sq = (xs[..., None] - _f(spec["modes"])) ** 2             # :497
    t = F32(-0.5) * sq - F32(spec["log_sqrt_2pi"]) + _f(spec["log_weights"])  # :500
    per_dim = _logsumexp_last(t.astype(np.float32))           # :504
    out = np.sum(per_dim, axis=-1, dtype=np.float32)          # :508
    return (out + log_jac).astype(np.float32)                 # :510


def logp_three_mixture(x, spec):
    """`ThreeMixtureDistributionTorch.log_density`, multimodal_torch.py:173-242 (batch branches)."""
    x = _f(x)
    means = _f(spec["means"])
    lw = _f(spec["log_weights"])
    c1 = _f(spec["c1"])
    comps = []
    if spec.get("scaling") is not None:                       # :200-212
        xs = x * _f(spec["scaling"])
        for k in range(3):
            c = xs - means[k]
            q = np.sum(c * c, axis=-1, dtype=np.float32)
            comps.append(((F32(-0.5) * q + c1[k]) + F32(spec["log_jacobian"])) + lw[k])
    else:                                                      # :227-242 (cov_inv = I)
        for k in range(3):
            c = x - means[k]
            q = np.sum(c * c, axis=-1, dtype=np.float32)
            comps.append((F32(-0.5) * q + c1[k]) + lw[k])
    t = np.stack(comps, axis=-1).astype(np.float32)
    return _logsumexp_last(t)


def logp_full_rosenbrock(x, spec):
    """`FullRosenbrockTorch.log_density`, target_distributions/rosenbrock_torch.py:67-84."""
    x = _f(x)
    a, b = F32(spec["a"]), F32(spec["b"])
    xi, xn = x[..., :-1], x[..., 1:]
    t1 = b * (xn - xi ** 2) ** 2                               # :75
    t2 = a * (xi - _f(spec["mu"])) ** 2                        # :76
    return (-(np.sum(t1, axis=-1, dtype=np.float32) + np.sum(t2, axis=-1, dtype=np.float32))).astype(np.float32)


def logp_even_rosenbrock(x, spec):
    """`EvenRosenbrockTorch.log_density`, rosenbrock_torch.py:194-210."""
    x = _f(x)
    a, b = F32(spec["a"]), F32(spec["b"])
    xo, xe = x[..., 0::2], x[..., 1::2]
    t1 = a * (xo - _f(spec["mu"])) ** 2                        # :203
    t2 = b * (xe - xo ** 2) ** 2                               # :204
    return (-(np.sum(t1, axis=-1, dtype=np.float32) + np.sum(t2, axis=-1, dtype=np.float32))).astype(np.float32)


def logp_hybrid_rosenbrock(x, spec):
    """`HybridRosenbrockTorch.log_density`, rosenbrock_torch.py:312-351."""
    x = _f(x)
    a, b, mu = F32(spec["a"]), F32(spec["b"]), F32(spec["mu"])
    n1, n2 = int(spec["n1"]), int(spec["n2"])
    x0 = x[..., 0]
    lp = -a * (x0 - mu) ** 2                                   # :319
    if x.shape[-1] > 1:
        blocks = x[..., 1:].reshape(x.shape[:-1] + (n2, n1 - 1))
        lp = lp - np.sum(b * (blocks[..., 0] - (x0 ** 2)[..., None]) ** 2, axis=-1, dtype=np.float32)  # :331-332
        for k in range(n1 - 2):                                # :337-345
            prev_sq = blocks[..., k] ** 2
            cur = blocks[..., k + 1]
            lp = lp - np.sum(b * (cur - prev_sq) ** 2, axis=-1, dtype=np.float32)
    return lp.astype(np.float32)


def logp_neal_funnel(x, spec):
    """`NealFunnelTorch.log_density`, target_distributions/funnel_torch.py:39-76."""
    x = _f(x)
    v = x[..., 0]
    l2p, lsv = F32(spec["log_2pi"]), F32(spec["log_sigma_v_sq"])
    mu_v, sv, mu_z, dm1 = F32(spec["mu_v"]), F32(spec["sigma_v_sq"]), F32(spec["mu_z"]), F32(spec["dm1"])
    prior = F32(-0.5) * l2p - F32(0.5) * lsv - F32(0.5) * (v - mu_v) ** 2 / sv   # :55
    if x.shape[-1] == 1:
        return prior.astype(np.float32)
    q = np.sum((x[..., 1:] - mu_z) ** 2, axis=-1, dtype=np.float32)               # :61
    with np.errstate(over="ignore", invalid="ignore"):
        lik = F32(-0.5) * dm1 * l2p - F32(0.5) * dm1 * v - F32(0.5) * np.exp(-v) * q  # :67-69
    return (prior + lik).astype(np.float32)


def logp_hypercube(x, spec):
    """`HypercubeTorch.log_density`, target_distributions/hypercube_torch.py:49-80."""
    x = _f(x)
    inside = np.all((x >= F32(spec["left"])) & (x <= F32(spec["right"])), axis=-1)
    return np.where(inside, F32(spec["log_uniform_density"]), _NEG_INF).astype(np.float32)


def logp_iid_gamma(x, spec):
    """`IIDGammaTorch.log_density`, target_distributions/iid_product_torch.py:52-91."""
    x = _f(x)
    k, th = F32(spec["shape"]), F32(spec["scale"])
    valid = np.all(x > 0, axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        body = np.sum((k - F32(1.0)) * np.log(x) - x / th, axis=-1, dtype=np.float32) - F32(spec["log_norm_const"])
    return np.where(valid, body, _NEG_INF).astype(np.float32)


def logp_iid_beta(x, spec):
    """`IIDBetaTorch.log_density`, iid_product_torch.py:188-229."""
    x = _f(x)
    al, be = F32(spec["alpha"]), F32(spec["beta"])
    valid = np.all((x > 0) & (x < 1), axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        body = np.sum((al - F32(1.0)) * np.log(x) + (be - F32(1.0)) * np.log(F32(1.0) - x),
                      axis=-1, dtype=np.float32) + F32(spec["log_norm_const"])
    return np.where(valid, body, _NEG_INF).astype(np.float32)


def logp_scaled_mvn(x, spec):
    """`ScaledMultivariateNormalTorch.log_density`, target_distributions/multivariate_normal_torch.py:198-223."""
    x = _f(x)
    sx = _f(spec["c"]) * x
    return (F32(spec["log_norm_const"]) - F32(0.5) * np.sum(sx ** 2, axis=-1, dtype=np.float32)).astype(np.float32)


def logp_mvn_diag(x, spec):
    """`MultivariateNormalTorch.log_density` (multivariate_normal_torch.py:62-92) for a DIAGONAL
    covariance: `(c @ cov_inv) * c` summed, where every off-diagonal product is an exact zero."""
    x = _f(x)
    c = x - _f(spec["mean"])
    q = np.sum((c * _f(spec["prec"])) * c, axis=-1, dtype=np.float32)
    return (F32(-0.5) * q + F32(spec["log_norm_const"])).astype(np.float32)


LOGP = {
    "rough_carpet": logp_rough_carpet,
    "three_mixture": logp_three_mixture,
    "full_rosenbrock": logp_full_rosenbrock,
    "even_rosenbrock": logp_even_rosenbrock,
    "hybrid_rosenbrock": logp_hybrid_rosenbrock,
    "neal_funnel": logp_neal_funnel,
    "hypercube": logp_hypercube,
    "iid_gamma": logp_iid_gamma,
    "iid_beta": logp_iid_beta,
    "scaled_mvn": logp_scaled_mvn,
    "mvn_diag": logp_mvn_diag,
}


def logp_mvn_dense(x, spec):
    """`MultivariateNormalTorch.log_density` with a general covariance, multivariate_normal_torch.py:62-92:
    temp = centered @ cov_inv; q = sum(temp * centered, dim=1); -0.5 q + log_norm_const (fp32; the matmul's summation
    order is the BLAS's, so comparisons are to ~1e-5 relative)."""
    cen = (x - _f(spec["mean"])[None, :]).astype(np.float32)
    temp = (cen @ _f(spec["cov_inv"])).astype(np.float32)
    q = np.sum(temp * cen, axis=1, dtype=np.float32)
    return (F32(-0.5) * q + F32(spec["log_norm_const"])).astype(np.float32)


def _log_sigmoid(z):
    z = z.astype(np.float32)
    return (np.minimum(z, F32(0.0)) - np.log1p(np.exp(-np.abs(z)))).astype(np.float32)


def logp_super_funnel(x, spec):
    """`SuperFunnelTorch.log_density`, funnel_torch.py:193-291: Bernoulli-logit likelihood over the N records
    (group, y, x[K]), Normal priors on alpha_j / beta_jk around the hyper-means with scales tau_alpha / tau_beta, Normal
    priors on the hyper-means, half-Cauchy priors on the taus; -inf when a tau is <= 1e-9."""
    J, K = int(spec["J"]), int(spec["K"])
    rec = _f(spec["records"])
    B = x.shape[0]
    al = x[:, :J]
    be = x[:, J:J + J * K].reshape(B, J, K)
    o = J + J * K
    mu_a, mu_b, tau_a, tau_b = x[:, o], x[:, o + 1:o + 1 + K], x[:, o + 1 + K], x[:, o + 2 + K]
    valid = (tau_a > 1e-9) & (tau_b > 1e-9)
    g = rec[:, 0].astype(np.int64)
    y = rec[:, 1][None, :]
    eta = (al[:, g] + np.einsum('nk,bnk->bn', rec[:, 2:], be[:, g, :])).astype(np.float32)
    ll = np.sum(y * _log_sigmoid(eta) + (F32(1.0) - y) * _log_sigmoid(-eta), axis=1, dtype=np.float32)
    sa = np.where(valid, tau_a, F32(1.0)).astype(np.float32)
    sb = np.where(valid, tau_b, F32(1.0)).astype(np.float32)
    l2p = F32(spec["log_2pi"])
    pa = np.sum(F32(-0.5) * l2p - np.log(sa)[:, None] - F32(0.5) * (al - mu_a[:, None]) ** 2 / (sa[:, None] ** 2), axis=1, dtype=np.float32)
    sq = np.sum((be - mu_b[:, None, :]) ** 2, axis=2, dtype=np.float32)
    pb = np.sum(F32(-0.5) * F32(K) * l2p - F32(K) * np.log(sb)[:, None] - F32(0.5) * sq / (sb[:, None] ** 2), axis=1, dtype=np.float32)
    hv, lhv = F32(spec["hyper_var"]), F32(spec["log_hyper_var"])
    p_ma = F32(-0.5) * l2p - F32(0.5) * lhv - F32(0.5) * mu_a ** 2 / hv
    p_mb = F32(-0.5) * F32(K) * l2p - F32(0.5) * F32(K) * lhv - F32(0.5) * np.sum(mu_b ** 2, axis=1, dtype=np.float32) / hv
    ts, lts = F32(spec["tau_scale"]), F32(spec["log_tau_scale"])
    with np.errstate(invalid="ignore"):
        p_ta = F32(spec["log_2"]) - F32(spec["log_pi"]) - lts - np.log1p((tau_a / ts) ** 2)
        p_tb = F32(spec["log_2"]) - F32(spec["log_pi"]) - lts - np.log1p((tau_b / ts) ** 2)
        out = (ll + pa + pb + p_ma + p_mb + p_ta + p_tb).astype(np.float32)
    return np.where(valid, out, F32(-np.inf)).astype(np.float32)


LOGP["mvn_dense"] = logp_mvn_dense
LOGP["super_funnel"] = logp_super_funnel


def log_density(spec: Dict, x) -> np.ndarray:
    """Dispatch on spec['family']; x is (d,) or (..., d)."""
    x = _f(x)
    single = x.ndim == 1
    out = LOGP[spec["family"]](x[None] if single else x, spec)
    return out[0] if single else out


# --------------------------------------------------------------------------------------
# Proposal transforms (raw randoms -> increments), temperature-aware scaling
# --------------------------------------------------------------------------------------

def normal_std(var: float, beta: float) -> np.float32:
    """`NormalProposal.__init__`, proposal_distributions/normal.py:27-31: sqrt(fp32(var/beta))."""
    return np.sqrt(F32(float(var) / float(beta)))


def normal_increments(z, var: float, beta: float):
    """`_sample_normal_jit`, normal.py:47-55: randn * std_dev."""
    return (_f(z) * normal_std(var, beta)).astype(np.float32)


def laplace_scale(var_vec, beta: float):
    """`LaplaceProposal.__init__`, proposal_distributions/laplace.py:29-32: sqrt((var_i/beta)/2) in fp32."""
    return np.sqrt((_f(var_vec) / F32(beta)) / F32(2.0)).astype(np.float32)


def laplace_increments(u01, var_vec, beta: float):
    """`_sample_laplace_jit`, laplace.py:47-69. u01 is the raw torch.rand draw in [0,1)."""
    u = _f(u01) - F32(0.5)
    clamped = np.maximum(F32(-2.0) * np.abs(u), F32(-0.999999))
    return (-laplace_scale(var_vec, beta)[None, ...] * np.sign(u) * np.log1p(clamped)).astype(np.float32)


def uniform_radius(radius: float, beta: float) -> np.float32:
    """`UniformRadiusProposal.__init__`, proposal_distributions/uniform.py:28-32: r / sqrt(fp32(beta))."""
    return F32(F32(radius) / np.sqrt(F32(beta)))


def uniform_radius_increments(z, u, radius: float, beta: float):
    """`_sample_uniform_ball_jit`, uniform.py:48-73: z/||z|| * R * u**(1/d)."""
    z = _f(z)
    d = z.shape[-1]
    norms = np.sqrt(np.sum(z * z, axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    safe = np.where(norms > F32(1e-12), norms, F32(1.0))
    dirs = z / safe
    radii = uniform_radius(radius, beta) * np.power(_f(u).reshape(z.shape[:-1] + (1,)), F32(1.0 / d))
    return (dirs * radii).astype(np.float32)


# --------------------------------------------------------------------------------------
# Initial state rule and temperature ladder
# --------------------------------------------------------------------------------------

def initial_state(target_name: str, dim: int, rng: np.random.RandomState) -> np.ndarray:
    """`MHAlgorithm.__init__`, interfaces/metropolis.py:21-64 (first match wins; drawn from NumPy's
    global RNG *before* the harness seeds it, interfaces/simulation_gpu.py:143-148)."""
    if "Beta" in target_name:
        return rng.uniform(0.2, 0.8, size=dim).astype(np.float32)
    if "Gamma" in target_name:
        return 5 + 0.01 * rng.randn(dim)
    if "RoughCarpet" in target_name or "ThreeMixture" in target_name:
        return np.zeros(dim)
    return 0.00000001 * rng.randn(dim)


def geometric_ladder() -> list:
    """`_construct_geometric_ladder`, algorithms/pt_rwm_gpu_optimized.py:245-257."""
    beta, ladder = 1.0, []
    while beta > 1e-2:
        ladder.append(beta)
        beta = beta * 0.5
    ladder.append(1e-2)
    return ladder


# --------------------------------------------------------------------------------------
# Accept rule, RWM run, PT run
# --------------------------------------------------------------------------------------

def accept_rule(lp_cur, lp_prop, u, beta):
    """`ultra_fused_mcmc_step_basic`, algorithms/rwm_gpu_optimized.py:21-25 (identical in
    `ultra_fused_parallel_mcmc_step`, pt_rwm_gpu_optimized.py:74-77):
    lar = beta*(lp' - lp); accept = (lar > 0) | (u < exp(lar)).  NaN -> reject."""
    with np.errstate(invalid="ignore", over="ignore"):
        lar = (_f(beta) * (_f(lp_prop) - _f(lp_cur))).astype(np.float32)
        acc = (lar > 0) | (_f(u) < np.exp(lar))
    return acc, lar


def rwm_run(spec: Dict, x0, beta, increments, uniforms, burn_in: int = 0,
            keep_states: bool = True, lp0=None) -> Dict:
    """B independent RWM chains with injected randomness.

    Follows `RandomWalkMH_GPU_Optimized.generate_samples` / `_single_step_ultra_fused`
    (algorithms/rwm_gpu_optimized.py:402-488, 289-336): proposal = x + inc (:310);
    lp' = log_density(proposal) (:311); accept rule (:314); acceptances counted only for
    steps > burn_in (:328-334); state t stored at chain index t (index 0 = x0, :376-380).

    x0 (B,d); beta scalar or (B,); increments (T,B,d); uniforms (T,B).
    Returns decisions (T,B) uint8, lar (T,B), chain (T+1,B,d) if keep_states, logp (T+1,B),
    accept_count (B,), acceptance_rate (B,), esjd (B,) per `expected_squared_jump_distance_gpu`.
    """
    x = _f(x0).copy()
    B, d = x.shape
    inc = _f(increments)
    us = _f(uniforms)
    T = inc.shape[0]
    beta = np.broadcast_to(_f(beta), (B,)).astype(np.float32)
    lp = log_density(spec, x) if lp0 is None else _f(lp0).copy()
    decisions = np.zeros((T, B), dtype=np.uint8)
    lars = np.zeros((T, B), dtype=np.float32)
    logps = np.zeros((T + 1, B), dtype=np.float32)
    logps[0] = lp
    chain = np.zeros((T + 1, B, d), dtype=np.float32) if keep_states else None
    if keep_states:
        chain[0] = x
    acc_count = np.zeros(B, dtype=np.int64)
    sq_sum = np.zeros(B, dtype=np.float64)
    for t in range(T):
        prop = (x + inc[t]).astype(np.float32)
        lpp = log_density(spec, prop)
        acc, lar = accept_rule(lp, lpp, us[t], beta)
        x_new = np.where(acc[:, None], prop, x).astype(np.float32)
        lp = np.where(acc, lpp, lp).astype(np.float32)
        if t + 1 > burn_in:
            acc_count += acc
            diff = (x_new - x).astype(np.float32)            # chain[t+1]-chain[t], :531
            sq_sum += np.sum(diff * diff, axis=-1, dtype=np.float32)
        x = x_new
        decisions[t] = acc
        lars[t] = lar
        logps[t + 1] = lp
        if keep_states:
            chain[t + 1] = x
    n_post = max(T - burn_in, 0)
    return {
        "decisions": decisions, "lar": lars, "chain": chain, "logp": logps,
        "final_state": x, "final_logp": lp, "accept_count": acc_count,
        "acceptance_rate": acc_count / max(n_post, 1),
        "esjd": sq_sum / max(n_post, 1),
    }


def swap_log_prob(beta_j, beta_k, lp_j, lp_k):
    """`fused_swap_probability_calculation`, algorithms/pt_rwm_gpu_optimized.py:42-47, literal order."""
    bj, bk, lj, lk = _f(beta_j), _f(beta_k), _f(lp_j), _f(lp_k)
    with np.errstate(invalid="ignore"):
        return (((bj * lk + bk * lj) - bj * lj) - bk * lk).astype(np.float32)


def pt_run(spec: Dict, x0, betas, increments, uniforms, swap_uniforms, swap_every: int,
           burn_in: int = 0, swap_mode: str = "reference", keep_states: bool = True) -> Dict:
    """L independent PT ladders of K temperatures with injected randomness.

    Follows `ParallelTemperingRWM_GPU_Optimized.step` / `_attempt_all_swaps`
    (algorithms/pt_rwm_gpu_optimized.py:541-574, 594-633): every step all K chains do a
    Metropolis step with their own beta (:546-567); when `s % swap_every == 0 and s > burn_in`
    (:544,570) a sequential sweep j = 0..K-2, k = j+1 draws one uniform per pair, computes the
    swap log-probability (:42-47), p = min(1, exp(.)) (:617) and accepts iff u < p (:621).
    swap_mode "reference": accepted pair copies state/logp of k into j and leaves k unchanged --
    what `fused_swap_execution_no_clone` (:50-59) does on tensor views.  swap_mode "exchange":
    textbook exchange (what the NumPy sampler does, algorithms/pt_rwm.py:137-155).
    After the sweep the K states are stored at chain index s (:574, 635-653).

    x0 (L,K,d); betas (K,) ; increments (T,L,K,d) post-scaling; uniforms (T,L,K);
    swap_uniforms (R,L,K-1) consumed one row per sweep.
    """
    x = _f(x0).copy()
    L, K, d = x.shape
    inc = _f(increments)
    us = _f(uniforms)
    su = _f(swap_uniforms)
    T = inc.shape[0]
    betas64 = np.asarray(betas, dtype=np.float64)   # the reference keeps the ladder as Python floats (:631)
    betas = _f(betas)                               # and the fp32 `beta_tensor` for the arithmetic (:216)
    lp = log_density(spec, x.reshape(L * K, d)).reshape(L, K)
    decisions = np.zeros((T, L, K), dtype=np.uint8)
    n_rounds = sum(1 for s in range(1, T + 1) if s % swap_every == 0 and s > burn_in)
    swap_dec = np.zeros((max(n_rounds, 1), L, max(K - 1, 1)), dtype=np.uint8)
    pre_sweep_logp = np.zeros((max(n_rounds, 1), L, K), dtype=np.float32)   # log-densities each sweep started from
    chain = np.zeros((T + 1, L, K, d), dtype=np.float32) if keep_states else None
    logps = np.zeros((T + 1, L, K), dtype=np.float32)
    logps[0] = lp
    if keep_states:
        chain[0] = x
    attempts = np.zeros(L, dtype=np.int64)
    accepts = np.zeros(L, dtype=np.int64)
    pair_accepts = np.zeros((L, max(K - 1, 1)), dtype=np.int64)
    attempts_at_last_accept = np.zeros(L, dtype=np.int64)
    sq_beta = np.zeros(L, dtype=np.float64)
    mh_accepts = np.zeros((L, K), dtype=np.int64)
    cold_sq = np.zeros((L, K), dtype=np.float64)
    r = 0
    for t in range(T):
        s = t + 1
        x_before = x
        prop = (x + inc[t]).astype(np.float32)
        lpp = log_density(spec, prop.reshape(L * K, d)).reshape(L, K)
        acc, _ = accept_rule(lp, lpp, us[t], betas[None, :])
        x = np.where(acc[..., None], prop, x).astype(np.float32)
        lp = np.where(acc, lpp, lp).astype(np.float32)
        decisions[t] = acc
        if s > burn_in:
            mh_accepts += acc
        if s % swap_every == 0 and s > burn_in and K > 1:
            x = x.copy()
            lp = lp.copy()
            pre_sweep_logp[r] = lp
            for j in range(K - 1):
                k = j + 1
                lsp = swap_log_prob(betas[j], betas[k], lp[:, j], lp[:, k])
                with np.errstate(over="ignore", invalid="ignore"):
                    p = np.minimum(F32(1.0), np.exp(lsp))
                ok = su[r, :, j] < p
                attempts += 1
                if swap_mode == "reference":
                    x[ok, j] = x[ok, k]
                    lp[ok, j] = lp[ok, k]
                else:
                    xj = x[ok, j].copy(); x[ok, j] = x[ok, k]; x[ok, k] = xj
                    lj = lp[ok, j].copy(); lp[ok, j] = lp[ok, k]; lp[ok, k] = lj
                accepts += ok
                pair_accepts[:, j] += ok
                attempts_at_last_accept = np.where(ok, attempts, attempts_at_last_accept)
                sq_beta += ok * (betas64[j] - betas64[k]) ** 2        # :631-632 (Python floats)
                swap_dec[r, :, j] = ok
            r += 1
        if s > burn_in:
            diff = (x - x_before).astype(np.float32)
            cold_sq += np.sum(diff * diff, axis=-1, dtype=np.float32)
        logps[t + 1] = lp
        if keep_states:
            chain[t + 1] = x
    n_post = max(T - burn_in, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref_rate = np.where(attempts_at_last_accept > 0, accepts / np.maximum(attempts_at_last_accept, 1), 0.0)
        ref_esjd = np.where(attempts_at_last_accept > 0, sq_beta / np.maximum(attempts_at_last_accept, 1), 0.0)
    return {
        "decisions": decisions, "swap_decisions": swap_dec[:max(n_rounds, 0)],
        "pre_sweep_logp": pre_sweep_logp[:max(n_rounds, 0)], "chain": chain, "logp": logps,
        "final_state": x, "final_logp": lp,
        "swap_attempts": attempts, "swap_accepts": accepts, "pair_accepts": pair_accepts,
        "attempts_at_last_accept": attempts_at_last_accept,
        # the reference only refreshes these two on an accepted swap (:627-633)
        "swap_acceptance_rate": ref_rate, "pt_esjd": ref_esjd,
        "sq_beta_jump_sum": sq_beta,
        "mh_accepts": mh_accepts,
        "esjd_per_temp": cold_sq / n_post,
        "cold_esjd": cold_sq[:, 0] / n_post,          # `expected_squared_jump_distance_gpu`, :772-789
    }


def esjd_from_chain(chain, burn_in: int) -> float:
    """`expected_squared_jump_distance_gpu`, algorithms/rwm_gpu_optimized.py:513-534: chain includes the
    initial state at index 0; uses chain[burn_in:], fp32 diffs, fp32 mean."""
    c = _f(chain)[burn_in:]
    if c.shape[0] < 2:
        return 0.0
    diff = c[1:] - c[:-1]
    sq = np.sum(diff * diff, axis=-1, dtype=np.float32)
    return float(np.mean(sq, dtype=np.float32))


# --------------------------------------------------------------------------------------
# The reference's NumPy CPU sampler (algorithms/rwm.py) restated, for the CPU-baseline leg
# --------------------------------------------------------------------------------------

def numpy_rwm_cpu(density_fn, dim: int, var: float, n_steps: int, seed: Optional[int], x0=None, beta: float = 1.0):
    """`RandomWalkMH.step` / `log_accept_prob`, algorithms/rwm.py:23-66, one chain, float64, NumPy global-RNG
    call order preserved (multivariate_normal then random): used to reproduce BASELINE.md section 3 row 1."""
    rs = np.random.RandomState(seed) if seed is not None else np.random
    x = np.zeros(dim) if x0 is None else np.asarray(x0, dtype=np.float64)
    cov = (var / beta) * np.eye(dim)
    lp = -np.inf
    chain = [x]
    n_acc = 0
    for _ in range(n_steps):
        prop = rs.multivariate_normal(x, cov)                   # :28
        dens = density_fn(prop)
        lpp = -np.inf if dens == 0 else math.log(dens + 1e-300)  # :52-55
        lar = beta * (lpp - lp)
        if lar > 0 or rs.random_sample() < math.exp(lar):       # :32
            x, lp = prop, lpp
            n_acc += 1
        chain.append(x)
    chain = np.asarray(chain)
    return {"chain": chain, "acceptance_rate": n_acc / len(chain),   # :36 denominator len(chain)
            "esjd": float(np.mean(np.sum((chain[1:] - chain[:-1]) ** 2, axis=1)))}


def rough_carpet_density_cpu(x, modes=(-15.0, 0.0, 15.0), weights=(0.5, 0.3, 0.2)) -> float:
    """`RoughCarpetDistribution.density`, target_distributions/multimodal.py:84-104 (unscaled)."""
    x = np.asarray(x, dtype=np.float64)
    c = 1.0 / math.sqrt(2 * math.pi)
    dens = np.zeros_like(x)
    for m, w in zip(modes, weights):
        dens = dens + w * np.exp(-0.5 * (x - m) ** 2) * c
    return float(np.prod(dens))


# -------------------------------------------------------------------------------------------------------------------
# Iterative temperature-ladder construction
# -------------------------------------------------------------------------------------------------------------------
def tempered_samples_dim(spec: Dict, d: int, n: int, beta: float, rs: np.random.RandomState) -> np.ndarray:
    """The targets' heuristic tempered samplers `draw_samples_torch(n, beta)`:
    RoughCarpet multimodal_torch.py:532-565 (mode per coordinate ~ weights, (m + z / sqrt(beta)) / s);
    ThreeMixture :270-333 (component per row, (mu_k + z / sqrt(beta)) / s); EvenRosenbrock rosenbrock_torch.py:224-248
    (x_{2j} ~ N(mu_j, 1 / (2 a beta)), x_{2j+1} | x_{2j} ~ N(x_{2j}^2, 1 / (2 b beta))); ScaledMVN
    multivariate_normal_torch.py:249-268; diagonal MVN :101-121."""
    fam = spec["family"]
    isb = F32(1.0) / np.sqrt(F32(beta))
    z = rs.standard_normal((n, d)).astype(np.float32)
    if fam == "rough_carpet":
        w = np.exp(_f(spec["log_weights"]).astype(np.float64)); w = w / w.sum()
        idx = rs.choice(3, size=(n, d), p=w)
        y = _f(spec["modes"])[idx] + z * isb
        return (y / _f(spec["scaling"])[None, :]).astype(np.float32) if "scaling" in spec else y.astype(np.float32)
    if fam == "three_mixture":
        w = np.exp(_f(spec["log_weights"]).astype(np.float64)); w = w / w.sum()
        idx = rs.choice(3, size=n, p=w)
        y = _f(spec["means"])[idx] + z * isb
        return (y / _f(spec["scaling"])[None, :]).astype(np.float32) if "scaling" in spec else y.astype(np.float32)
    if fam == "even_rosenbrock":
        a, b = float(spec["a"]) * beta, float(spec["b"]) * beta
        sa = np.sqrt(F32(1.0 / (2 * a))) if a > 0 else F32(1.0)
        sb = np.sqrt(F32(1.0 / (2 * b))) if b > 0 else F32(1.0)
        x = np.zeros((n, d), np.float32)
        x[:, 0::2] = _f(spec["mu"])[None, :] + z[:, 0::2] * sa
        x[:, 1::2] = x[:, 0::2] ** 2 + z[:, 1::2] * sb
        return x
    if fam == "scaled_mvn":
        return (z * isb / _f(spec["c"])[None, :]).astype(np.float32)
    if fam == "mvn_diag":
        return (_f(spec["mean"])[None, :] + z * isb / np.sqrt(_f(spec["prec"]))[None, :]).astype(np.float32)
    raise NotImplementedError(fam)


def swap_probability(spec: Dict, d: int, beta_curr: float, beta_star: float, n: int, rs: np.random.RandomState):
    """The estimate inside `_construct_iterative_ladder` (algorithms/pt_rwm_gpu_optimized.py:356-368):
    mean(exp(clamp_max((beta - beta*) (log pi(x*) - log pi(x)), 0))), x* ~ sampler(beta*), x ~ sampler(beta).
    Returns (estimate, Monte-Carlo standard error)."""
    xs = tempered_samples_dim(spec, d, n, beta_star, rs)
    xc = tempered_samples_dim(spec, d, n, beta_curr, rs)
    log_r = F32(beta_curr - beta_star) * (log_density(spec, xs) - log_density(spec, xc))
    p = np.exp(np.minimum(log_r, F32(0.0))).astype(np.float64)
    return float(p.mean()), float(p.std(ddof=1) / np.sqrt(n))


def iterative_ladder(estimate, target_rate: float = 0.234, beta_min: float = 0.01, n_samples: int = 3000, tolerance: float = 0.005,
                     initial_pn: float = 0.5, pn_update_power: float = -0.25, max_pn_steps: int = 100,
                     pn_clamp=(-10.0, 10.0), fail_tol_factor: float = 3.0) -> list:
    """`_construct_iterative_ladder` (algorithms/pt_rwm_gpu_optimized.py:283-426) with a pluggable estimator
    `estimate(beta_curr, beta_star, n) -> float`: beta* = beta / (1 + exp(p_n)); p_n <- p_n + n^power (a - target) until
    |a - target| <= tolerance; a rung is also accepted after max_pn_steps when within fail_tol_factor * tolerance; the
    ladder is closed with beta_min."""
    ladder, beta_curr = [1.0], 1.0
    while True:
        if beta_curr <= beta_min + 1e-6:
            break
        pn, n_upd, found = initial_pn, 1, False
        last_star, last_prob, it = -1.0, -1.0, 0
        for it in range(1, max_pn_steps + 1):
            cpn = float(np.clip(pn, pn_clamp[0], pn_clamp[1]))
            beta_star = beta_curr / (1.0 + np.exp(cpn))
            last_star = beta_star
            if beta_star < beta_min:
                break
            prob = estimate(beta_curr, beta_star, n_samples)
            last_prob = prob
            if abs(prob - target_rate) <= tolerance:
                ladder.append(beta_star)
                beta_curr, found = beta_star, True
                break
            pn = pn + (n_upd ** pn_update_power) * (prob - target_rate)
            n_upd += 1
        if not found:
            if it == max_pn_steps and last_star >= beta_min and last_star != -1.0 and \
                    abs(last_prob - target_rate) <= tolerance * fail_tol_factor:
                ladder.append(last_star)
                beta_curr = last_star
            else:
                break
    if ladder[-1] > beta_min + 1e-5:
        ladder.append(beta_min)
    return ladder
