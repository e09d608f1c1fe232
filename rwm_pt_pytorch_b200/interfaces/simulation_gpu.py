"""`MCMCSimulation_GPU` -- the harness the reference's drivers use (interfaces/simulation_gpu.py:13-437).

Same constructor, the same class-name dispatch ('GPU' / 'ParallelTempering' in `algorithm.__name__`, :81-83), the same
late seeding (:143-148, which here really seeds the run because the Philox key is drawn from torch's generator when
sampling starts), the same delegation to `algorithm.generate_samples` and the same accessors.  Plotting
(`traceplot`, `samples_histogram`) needs matplotlib and is out of scope."""
from __future__ import annotations

import time
from typing import Optional, Union

import numpy as np
import torch

from .target import TargetDistribution
from .target_torch import TorchTargetDistribution
from ..proposal_distributions import ProposalDistribution, NormalProposal, LaplaceProposal, UniformRadiusProposal


class MCMCSimulation_GPU:
    def __init__(self, dim: int, sigma: float = None, proposal_config: dict = None, num_iterations: int = 1000,
                 algorithm=None, target_dist: Union[TargetDistribution, TorchTargetDistribution] = None,
                 symmetric: bool = True, seed: Optional[int] = None, beta_ladder: Optional[list] = None,
                 swap_acceptance_rate: Optional[float] = None, device: Optional[str] = None, pre_allocate: bool = True,
                 burn_in: int = 0, **kwargs):
        if proposal_config is None and sigma is not None:
            proposal_config = {'name': 'Normal', 'params': {'base_variance_scalar': sigma}}
        elif proposal_config is None and sigma is None:
            raise ValueError("Either sigma (backward compatibility) or proposal_config must be provided")
        self.num_iterations = num_iterations
        self.burn_in = max(0, burn_in)
        self.target_dist = target_dist
        self.proposal_config = proposal_config
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'
        self.device = device
        self.pre_allocate = pre_allocate
        name = getattr(algorithm, '__name__', '')
        if 'GPU' in name and 'ParallelTempering' in name:
            pt_kwargs = dict(kwargs)
            if swap_acceptance_rate is not None:
                pt_kwargs['swap_acceptance_rate'] = swap_acceptance_rate
            self.algorithm = algorithm(dim, sigma, target_dist, symmetric, device=device,
                                       pre_allocate_steps=num_iterations if pre_allocate else None,
                                       beta_ladder=beta_ladder, burn_in=self.burn_in, **pt_kwargs)
        elif 'GPU' in name:
            algo_beta = beta_ladder[0] if beta_ladder else 1.0
            proposal_dist = self._create_proposal_distribution(
                dim=dim, beta=algo_beta, proposal_config=proposal_config, device=torch.device(device),
                dtype=torch.float32, use_efficient_rng=kwargs.get('use_efficient_rng', True))
            self.algorithm = algorithm(
                dim=dim, proposal_distribution=proposal_dist, target_dist=target_dist, symmetric=symmetric,
                beta=algo_beta, device=device, pre_allocate_steps=num_iterations if pre_allocate else None,
                burn_in=self.burn_in, use_efficient_rng=kwargs.get('use_efficient_rng', True),
                **{k: v for k, v in kwargs.items() if k != 'use_efficient_rng'})
        else:
            raise TypeError("MCMCSimulation_GPU drives the GPU samplers of this package "
                            "(class name must contain 'GPU'); the NumPy CPU samplers are not part of it")
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed(seed)
        self._start_time = None
        self._end_time = None

    def reset(self):
        self.algorithm.reset()

    def has_run(self):
        steps = getattr(self.algorithm, 'total_steps', None)
        if steps is None:
            steps = getattr(self.algorithm, 'step_counter', 0)
        return steps > 0

    def generate_samples(self, progress_bar=True, as_list: bool = True):
        """Delegates to the algorithm's single-launch `generate_samples` (:165-212).  `as_list=True` keeps the
        reference's return type (nested Python lists, :189-190); pass False to get the device tensor."""
        if self.has_run():
            raise ValueError("Please reset the algorithm before running it again.")
        self._start_time = time.time()
        chain = self.algorithm.generate_samples(self.num_iterations)
        if as_list and hasattr(chain, 'cpu'):
            chain = chain.cpu().numpy().tolist()
        self._end_time = time.time()
        return chain

    def acceptance_rate(self):
        if not self.has_run():
            raise ValueError("The algorithm has not been run yet.")
        return self.algorithm.acceptance_rate

    def expected_squared_jump_distance(self):
        if not self.has_run():
            raise ValueError("The algorithm has not been run yet.")
        return self.algorithm.expected_squared_jump_distance_gpu()

    def pt_expected_squared_jump_distance(self):
        if not self.has_run():
            raise ValueError("The algorithm has not been run yet.")
        return self.algorithm.pt_esjd

    def benchmark_performance(self, num_samples_list=(1000, 5000, 10000, 50000), compare_cpu=True):
        """Wall-clock samples/s for several run lengths (:252-311).  As in the reference the "CPU" arm re-runs the SAME
        algorithm object after a reset (:291-304) -- it is a second timing of the GPU path, kept for schema compatibility;
        the host-CPU baselines of this repository are in bench.py."""
        results = {'sample_sizes': list(num_samples_list), 'gpu_times': [], 'gpu_samples_per_sec': [],
                   'cpu_times': [] if compare_cpu else None, 'cpu_samples_per_sec': [] if compare_cpu else None,
                   'speedup': [] if compare_cpu else None}

        def timed(n):
            self.reset()
            self.num_iterations = n
            t0 = time.time()
            self.generate_samples(progress_bar=False, as_list=False)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            return max(time.time() - t0, 1e-9)

        original = self.num_iterations
        try:
            for n in num_samples_list:
                dt = timed(n)
                results['gpu_times'].append(dt)
                results['gpu_samples_per_sec'].append(n / dt)
                if compare_cpu:
                    dt2 = timed(n)
                    results['cpu_times'].append(dt2)
                    results['cpu_samples_per_sec'].append(n / dt2)
                    results['speedup'].append(dt2 / dt)
        finally:
            self.num_iterations = original
        return results

    def traceplot(self, *a, **k):
        raise NotImplementedError("plotting is out of scope of the sm_100a sampling path (matplotlib not bundled)")

    samples_histogram = traceplot

    def _create_proposal_distribution(self, dim: int, beta: float, proposal_config: dict, device: torch.device,
                                      dtype: torch.dtype, use_efficient_rng: bool = True) -> ProposalDistribution:
        """Proposal plugin from `{'name': ..., 'params': {...}}` (:380-437); same errors as the reference."""
        name = proposal_config.get('name')
        params = proposal_config.get('params', {})
        if name == "Normal":
            v = params.get('base_variance_scalar')
            if v is None:
                raise ValueError("Normal proposal requires 'base_variance_scalar' parameter")
            return NormalProposal(dim, v, beta, device, dtype, None)
        if name == "Laplace":
            v = params.get('base_variance_vector')
            if v is None:
                raise ValueError("Laplace proposal requires 'base_variance_vector' parameter")
            if isinstance(v, (list, tuple)):
                v = torch.tensor(v, dtype=dtype)
            elif isinstance(v, (int, float)):
                v = torch.full((dim,), float(v), dtype=dtype)
            elif isinstance(v, torch.Tensor):
                v = v.to(dtype=dtype)
            else:
                raise ValueError(f"Invalid base_variance_vector type: {type(v)}")
            return LaplaceProposal(dim, v, beta, device, dtype, None)
        if name == "UniformRadius":
            r = params.get('base_radius')
            if r is None:
                raise ValueError("UniformRadius proposal requires 'base_radius' parameter")
            return UniformRadiusProposal(dim, r, beta, device, dtype, None)
        raise ValueError(f"Unknown proposal distribution name: {name}")
