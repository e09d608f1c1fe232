"""Sweep drivers of the reference, batched.

`run_rwm_study` is the 40-scale RWM sweep of experiment_RWM_GPU.py:165-301 -- but all scale values (times
`chains_per_value` independent chains each) run in ONE launch of the fused kernel, each chain with its own proposal
scale -- and `run_pt_study` is the 30-target-swap-rate PT sweep of experiment_pt_GPU.py:165-279 (one launch per rate:
every rate builds its own iterative ladder).  Both write the reference's JSON schema, so `data/average_seeds.py` and
`plot.py` of the reference keep working on the files; extra keys are appended only."""
from __future__ import annotations

import json
import os
import time
from typing import Optional

import numpy as np
import torch

from . import target_distributions as td
from .algorithms import RandomWalkMH_GPU_Optimized, ParallelTemperingRWM_GPU_Optimized
from .proposal_distributions import NormalProposal, LaplaceProposal, UniformRadiusProposal


def get_target_distribution(name: str, dim: int, device=None, pt: bool = False, **kwargs):
    """Target factory with the experiment defaults of the reference: experiment_RWM_GPU.py:21-117 (RoughCarpet modes +-4,
    ThreeMixture centres +-5) or, with pt=True, experiment_pt_GPU.py:21-117 (modes / centres +-15)."""
    if device is None:
        device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    sep = 15.0 if pt else 5.0
    three_centers = [[-sep] + [0.0] * (dim - 1), [0.0] * dim, [sep] + [0.0] * (dim - 1)]
    if pt:
        kwargs.setdefault('mode_centers', three_centers if name.startswith("ThreeMixture") else [-15.0, 0.0, 15.0])
    if name == "MultivariateNormal":
        return td.MultivariateNormalTorch(dim, device=device)
    if name == "MultivariateNormalScaled":
        return td.ScaledMultivariateNormalTorch(dim, device=device)
    if name in ("RoughCarpet", "RoughCarpetScaled"):
        return td.RoughCarpetDistributionTorch(dim, scaling=name.endswith("Scaled"), device=device,
                                               mode_centers=kwargs.get('mode_centers', [-4.0, 0.0, 4.0]),
                                               mode_weights=kwargs.get('mode_weights', [0.5, 0.3, 0.2]))
    if name in ("ThreeMixture", "ThreeMixtureScaled"):
        return td.ThreeMixtureDistributionTorch(dim, scaling=name.endswith("Scaled"), device=device,
                                                mode_centers=kwargs.get('mode_centers', three_centers),
                                                mode_weights=kwargs.get('mode_weights', [1 / 3, 1 / 3, 1 / 3]))
    if name == "Hypercube":
        return td.HypercubeTorch(dim, left_boundary=-1, right_boundary=1, device=device)
    if name == "IIDGamma":
        return td.IIDGammaTorch(dim, shape=2, scale=3, device=device)
    if name == "IIDBeta":
        return td.IIDBetaTorch(dim, alpha=2, beta=3, device=device)
    rb = dict(a_coeff=kwargs.get('a_coeff', 1.0 / 20.0), b_coeff=kwargs.get('b_coeff', 100.0 / 20.0), mu=kwargs.get('mu', 1.0))
    if name == "FullRosenbrock":
        return td.FullRosenbrockTorch(dim, device=device, **rb)
    if name == "EvenRosenbrock":
        return td.EvenRosenbrockTorch(dim, device=device, **rb)
    if name == "HybridRosenbrock":
        return td.HybridRosenbrockTorch(n1=kwargs.get('n1', 3), n2=kwargs.get('n2', 5), device=device, **rb)
    if name == "NealFunnel":
        return td.NealFunnelTorch(dim, mu_v=kwargs.get('mu_v', 0.0), sigma_v_sq=kwargs.get('sigma_v_sq', 9.0),
                                  mu_z=kwargs.get('mu_z', 0.0), device=device)
    if name == "SuperFunnel":
        # synthetic data exactly as the reference generates it (experiment_RWM_GPU.py:95-120), on the CPU generator so that
        # the data do not depend on the device
        J, K, n_per = kwargs.get('J', 5), kwargs.get('K', 3), kwargs.get('n_per_group', 20)
        torch.manual_seed(42)
        X_data, Y_data = [], []
        for _ in range(J):
            X_j = torch.randn(n_per, K)
            Y_j = torch.bernoulli(torch.sigmoid(0.5 * torch.sum(X_j, dim=1)))
            X_data.append(X_j)
            Y_data.append(Y_j)
        return td.SuperFunnelTorch(J, K, X_data, Y_data, prior_hypermean_std=kwargs.get('prior_hypermean_std', 10.0),
                                   prior_tau_scale=kwargs.get('prior_tau_scale', 2.5), device=device)
    raise ValueError("Unknown target distribution name")


def _se(x: np.ndarray) -> np.ndarray:
    return x.std(axis=1, ddof=1) / np.sqrt(x.shape[1]) if x.shape[1] > 1 else np.zeros(x.shape[0])


def run_rwm_study(dim, target_name="MultivariateNormal", num_iters=100000, var_max=3.5, seed=42, burn_in=1000,
                  proposal_name="Normal", proposal_params=None, num_values: int = 40, chains_per_value: int = 1,
                  out_dir: Optional[str] = None, device=None, **kwargs) -> dict:
    """ESJD / acceptance versus proposal scale (the 0.234 study): `num_values` scales in linspace(0.01, var_max) --
    Normal / Laplace variance = scale^2 / dim, UniformRadius radius = scale (experiment_RWM_GPU.py:202-239) -- times
    `chains_per_value` chains, one kernel launch."""
    target = get_target_distribution(target_name, dim, device=device, **kwargs)
    d = target.dim
    scales = np.linspace(0.01, var_max, num_values)
    n_chains = num_values * chains_per_value
    cpu = torch.device("cpu")
    if proposal_name == "Normal":
        prop = NormalProposal(d, 1.0, 1.0, cpu, torch.float32)
        k_scales = np.sqrt((scales ** 2 / d).astype(np.float32))                  # std = sqrt(fp32(var))
    elif proposal_name == "Laplace":
        aniso = proposal_params.get('anisotropic') if proposal_params else None
        base = torch.tensor(aniso, dtype=torch.float32) if aniso is not None else torch.ones(d)
        prop = LaplaceProposal(d, base, 1.0, cpu, torch.float32)                 # scale_i = sqrt(base_i / 2)
        k_scales = np.sqrt((scales ** 2 / d).astype(np.float32))                  # times sqrt(effective variance)
    elif proposal_name == "UniformRadius":
        prop = UniformRadiusProposal(d, 1.0, 1.0, cpu, torch.float32)
        k_scales = scales.astype(np.float32)
    else:
        raise ValueError(f"Unknown proposal name: {proposal_name}")
    torch.manual_seed(seed)
    np.random.seed(seed)
    t0 = time.time()
    algo = RandomWalkMH_GPU_Optimized(d, target_dist=target, burn_in=burn_in, device=device or "cuda", proposal_distribution=prop,
                                      num_chains=n_chains, store="none", seed=seed,
                                      proposal_scales=np.repeat(k_scales, chains_per_value))
    algo.generate_samples(num_iters)
    acc = algo.acceptance_rates.cpu().numpy().reshape(num_values, chains_per_value)
    esjd = algo.esjd_per_chain().cpu().numpy().reshape(num_values, chains_per_value)
    total_time = time.time() - t0
    acceptance_rates, esjds = acc.mean(axis=1).tolist(), esjd.mean(axis=1).tolist()
    best = int(np.argmax(esjds))
    data = {
        'target_distribution': target_name, 'proposal_distribution': proposal_name, 'dimension': d,
        'num_iterations': num_iters, 'seed': seed, 'total_time': total_time, 'max_esjd': esjds[best],
        'max_acceptance_rate': acceptance_rates[best], 'max_scale_param': float(scales[best]),
        'expected_squared_jump_distances': esjds, 'acceptance_rates': acceptance_rates,
        'scale_param_range': scales.tolist(), 'times': [total_time / num_values] * num_values,
        # appended keys ('var_value_range' / 'max_variance_value': the names the reference's earlier files and its
        # data/average_seeds.py use for the same two fields)
        'var_value_range': scales.tolist(), 'max_variance_value': float(scales[best]),
        'chains_per_value': chains_per_value, 'acceptance_rate_se': _se(acc).tolist(), 'esjd_se': _se(esjd).tolist(),
        'burn_in': burn_in, 'chain_steps_per_sec': n_chains * (num_iters + burn_in) / total_time,
    }
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        fn = os.path.join(out_dir, f"{target_name}_{proposal_name}_RWM_GPU_dim{d}_{num_iters}iters_seed{seed}.json")
        with open(fn, "w") as f:
            json.dump(data, f, indent=2)
        data['filename'] = fn
    return data


def run_pt_study(dim, target_name="ThreeMixture", num_iters=100000, swap_accept_max=0.5, seed=42, burn_in=1000,
                 N_samples_swap_est=50000, iterative_tolerance=0.0005, iterative_max_pn_steps=500,
                 iterative_fail_tol_factor=1.5, num_values: int = 30, ladders_per_value: int = 1, swap_every: int = 100,
                 out_dir: Optional[str] = None, device=None, **kwargs) -> dict:
    """PT-ESJD versus target swap rate (experiment_pt_GPU.py:165-279): each of the `num_values` rates in
    linspace(0.01, swap_accept_max) builds its iterative ladder and runs `ladders_per_value` ladders in one launch."""
    target = get_target_distribution(target_name, dim, device=device or "cuda", pt=True, **kwargs)
    d = target.dim
    rates = np.linspace(0.01, swap_accept_max, num_values)
    var = (2.38 ** 2) / d
    acc, esjd, times, sizes = [], [], [], []
    t_all = time.time()
    for i, rate in enumerate(rates):
        t0 = time.time()
        torch.manual_seed(seed)
        np.random.seed(seed)
        algo = ParallelTemperingRWM_GPU_Optimized(
            d, var, target, swap_acceptance_rate=float(rate), iterative_temp_spacing=True, N_samples_swap_est=N_samples_swap_est,
            iterative_tolerance=iterative_tolerance, iterative_max_pn_steps=iterative_max_pn_steps,
            iterative_fail_tol_factor=iterative_fail_tol_factor, swap_every=swap_every, burn_in=burn_in, device=device or "cuda",
            num_ladders=ladders_per_value, store="none", seed=seed + i)
        algo.generate_samples(num_iters)
        acc.append(float(algo.swap_acceptance_rate))
        esjd.append(float(algo.pt_esjd))
        sizes.append(algo.num_chains)
        times.append(time.time() - t0)
    best = int(np.argmax(esjd))
    data = {
        'target_distribution': target_name, 'dimension': d, 'num_iterations': num_iters, 'seed': seed,
        'total_time': time.time() - t_all, 'max_esjd': esjd[best], 'max_actual_acceptance_rate': acc[best],
        'max_constr_acceptance_rate': float(rates[best]), 'expected_squared_jump_distances': esjd, 'acceptance_rates': acc,
        'swap_acceptance_rates_range': rates.tolist(), 'times': times,
        'ladder_sizes': sizes, 'ladders_per_value': ladders_per_value, 'burn_in': burn_in,
    }
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        fn = os.path.join(out_dir, f"{target_name}_PT_GPU_dim{d}_{num_iters}iters_seed{seed}.json")
        with open(fn, "w") as f:
            json.dump(data, f, indent=2)
        data['filename'] = fn
    return data
