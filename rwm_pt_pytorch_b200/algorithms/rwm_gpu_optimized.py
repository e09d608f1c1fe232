"""`RandomWalkMH_GPU_Optimized` -- drop-in for the reference class of the same name
(algorithms/rwm_gpu_optimized.py:79-579), backed by the persistent fused sm_100a kernel.

Same constructor arguments (same order), methods and attributes as the reference.  New optional keyword
arguments are appended at the end only: `num_chains` (batch of independent chains in one launch), `seed`,
`store`, `thin`, `math_mode`, `initial_states`, `chain_id_base`, `lanes_per_chain`, `proposal_scales`.

Differences that are deliberate: the Python per-step loop, the pre-generated (T, d) random
tensors and the per-step `.item()` sync are gone -- one kernel launch runs all burn_in + num_samples steps with
in-kernel Philox; `seed` (or `torch.manual_seed`) really seeds the stream; no TF32 switch is touched.
"""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..interfaces import MHAlgorithm, TargetDistribution, TorchTargetDistribution
from ..proposal_distributions import ProposalDistribution, NormalProposal, LaplaceProposal, UniformRadiusProposal
from ._engine import LadderBatch


class RandomWalkMH_GPU_Optimized(MHAlgorithm):
    """Random-Walk Metropolis for `num_chains` independent chains on one B200."""

    def __init__(self, dim: int,
                 var: float = None,
                 target_dist: TorchTargetDistribution | TargetDistribution = None,
                 symmetric: bool = True,
                 beta: float = 1.0,
                 burn_in: int = 0,
                 device: str = None,
                 pre_allocate_steps: int = None,
                 use_efficient_rng: bool = True,
                 compile_mode: str = None,
                 proposal_distribution: ProposalDistribution = None,
                 # ---- appended keyword arguments (not in the reference) ----
                 num_chains: int = 1,
                 seed: Optional[int] = None,
                 store: Optional[str] = None,
                 thin: int = 1,
                 math_mode: str = "fast",
                 initial_states=None,
                 chain_id_base: int = 0,
                 lanes_per_chain: int = 0,
                 proposal_scales=None):
        if not isinstance(target_dist, TorchTargetDistribution):
            raise TypeError("RandomWalkMH_GPU_Optimized needs a TorchTargetDistribution from "
                            "rwm_pt_pytorch_b200.target_distributions (the legacy NumPy-density path of the "
                            "reference, rwm_gpu_optimized.py:368-372, is a host loop and is not provided)")
        self.num_chains = int(num_chains)
        if self.num_chains < 1:
            raise ValueError("num_chains must be >= 1")
        var_arr = None
        if proposal_distribution is not None:
            super().__init__(dim, 1.0, target_dist, symmetric, num_chains=self.num_chains)  # nominal var (:118-120)
        elif var is not None:
            var_arr = np.asarray(var, dtype=np.float64)
            super().__init__(dim, var if var_arr.ndim == 0 else var_arr, target_dist, symmetric, num_chains=self.num_chains)
            if np.any(var_arr <= 0):
                raise ValueError("base_variance_scalar must be positive")
        else:
            raise ValueError("Either var (backward compatibility) or proposal_distribution must be provided")

        if device is None:
            self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        else:
            self.device = torch.device(device)
        self.dtype = torch.float32
        self.use_efficient_rng = use_efficient_rng
        self.compile_mode = compile_mode
        self.rng_generator = None  # kept for API compatibility; randomness is in-kernel Philox
        beta_arr = np.asarray(beta, dtype=np.float64)
        self.beta = beta
        self.beta_tensor = torch.tensor(beta_arr, dtype=torch.float32)

        if proposal_distribution is None:
            v0 = float(var_arr) if var_arr.ndim == 0 else float(var_arr.reshape(-1)[0])
            b0 = float(beta_arr) if beta_arr.ndim == 0 else float(beta_arr.reshape(-1)[0])
            proposal_distribution = NormalProposal(dim=dim, base_variance_scalar=v0, beta=b0,
                                                   device=torch.device("cpu"), dtype=torch.float32)
        self.proposal_dist = proposal_distribution
        self.name = f"RWM_GPU_FUSED_{self.proposal_dist.get_name()}"

        # per-chain proposal scale / beta (variance sweeps batch many `var` values into one launch)
        betas = np.broadcast_to(beta_arr, (self.num_chains,)).astype(np.float64)
        if var_arr is not None and isinstance(self.proposal_dist, NormalProposal):
            vars_ = np.broadcast_to(var_arr, (self.num_chains,)).astype(np.float64)
            scales = np.sqrt((vars_ / betas).astype(np.float32)).astype(np.float32)     # normal.py:27-31
        elif (beta_arr.ndim == 0 and torch.device(self.proposal_dist.device).type == self.device.type
              and self.proposal_dist.dtype == self.dtype):
            # a supplied proposal that already lives on the sampler's device / dtype is used AS IS by the reference, i.e.
            # with the proposal's own beta (rwm_gpu_optimized.py:209-212); only a proposal on another device / dtype is
            # rebuilt with the sampler's beta (:165-208)
            scales = np.full(self.num_chains, self.proposal_dist._sample_scale(), dtype=np.float32)
        else:
            scales = np.asarray([self.proposal_dist.chain_scale(float(b)) for b in betas], dtype=np.float32)
        if proposal_scales is not None:
            # explicit per-chain kernel scale (std / Laplace multiplier / ball radius): batches a whole scale sweep
            scales = np.broadcast_to(np.asarray(proposal_scales, dtype=np.float32), (self.num_chains,)).copy()
            if np.any(scales <= 0):
                raise ValueError("proposal_scales must be positive")
        self._betas, self._scales = betas.astype(np.float32), scales

        self.num_acceptances = 0
        self.acceptance_rate = 0.0
        self.total_steps = 0
        self.burn_in = max(0, burn_in)
        self.pre_allocate_steps = pre_allocate_steps
        self.thin = max(1, int(thin))
        if store is None:
            store = "all" if (self.num_chains == 1 or pre_allocate_steps) else "none"
        if store not in ("all", "none"):
            raise ValueError("store must be 'all' or 'none'")
        self.store = store
        self.math_mode = math_mode
        self.seed = seed
        self.chain_id_base = chain_id_base
        self.lanes_per_chain = lanes_per_chain
        if initial_states is not None:
            x0 = np.asarray(initial_states, dtype=np.float64).reshape(self.num_chains, dim)
            self._x0 = x0
            self.chain = [x0[0]]
        self.pre_allocated_chain = None
        self.pre_allocated_log_densities = None
        self.chain_index = 0 if pre_allocate_steps else None
        self.current_state = None
        self.log_target_density_current = None
        self.use_torch_target = True
        self.compiled_log_density = None
        self.precomputed_increments = None
        self.precomputed_random_vals = None
        self.increment_index = 0
        self._batch: Optional[LadderBatch] = None
        self.acceptance_rates = None

    # ------------------------------------------------------------------------------------------------
    def get_name(self):
        return self.name

    def reset(self):
        """Back to the initial state (reference: rwm_gpu_optimized.py:271-283)."""
        super().reset()
        self.num_acceptances = 0
        self.acceptance_rate = 0.0
        self.total_steps = 0
        self.current_state = None
        self.log_target_density_current = None
        self.pre_allocated_chain = None
        self.pre_allocated_log_densities = None
        self.chain_index = 0 if self.pre_allocate_steps else None
        self._batch = None
        self.acceptance_rates = None

    def _ensure_batch(self, rows_needed: int):
        if self._batch is None:
            self._batch = LadderBatch(
                self.target_dist, self.dim, self.num_chains, 1, self._betas[:, None], self.proposal_dist.family_id,
                self._scales[:, None], self.proposal_dist.dim_scale(), self._x0, self.device, burn_in=self.burn_in,
                math_mode=self.math_mode, seed=self.seed, chain_id_base=self.chain_id_base,
                lanes_per_chain=self.lanes_per_chain)
            if self.store == "all":
                cap = rows_needed
                if self.pre_allocate_steps:
                    cap = self.burn_in + self.pre_allocate_steps + 1       # :226-228
                self._batch.allocate_storage("all", cap, self.thin)
        elif self.store == "all" and rows_needed > self._batch.capacity:
            if self.pre_allocate_steps:
                warnings.warn("Pre-allocated chain full, switching to dynamic allocation")   # :382
            self._batch.grow_storage(rows_needed)
        self._sync_views()

    def _sync_views(self):
        b = self._batch
        self.current_state = b.state[0] if self.num_chains == 1 else b.state
        self.log_target_density_current = b.logp[0] if self.num_chains == 1 else b.logp
        if b.samples is not None:
            self.pre_allocated_chain = b.samples[0] if self.num_chains == 1 else b.samples
            self.pre_allocated_log_densities = b.sample_logp[0] if self.num_chains == 1 else b.sample_logp
            self.chain_index = b.rows_written()

    def _refresh_stats(self):
        b = self._batch
        self.total_steps = b.total_steps
        post = b.post_burn_in_steps()
        acc = b.accept_count
        self.num_acceptances = int(acc.sum().item())
        if post > 0:
            self.acceptance_rates = acc.to(torch.float64) / post
            self.acceptance_rate = self.num_acceptances / (post * self.num_chains)   # :332-334 (pooled over chains)
        self._sync_views()

    def step(self):
        """One Metropolis step for every chain (one launch of the fused kernel with n_steps = 1)."""
        self._ensure_batch(self.total_steps // self.thin + 2)
        self._batch.run(1)
        self._refresh_stats()

    def generate_samples(self, num_samples: int):
        """Run burn_in + num_samples steps in ONE kernel launch and return the retained post-burn-in samples:
        (num_samples, dim) for a single chain as the reference does (:482-488), (num_chains, num_samples, dim)
        for a batch, or an empty (num_chains, 0, dim) tensor when store='none' (accumulators only)."""
        total = self.burn_in + int(num_samples)          # :424 (every call runs burn_in + num_samples steps)
        first_row = 1 + self.burn_in // self.thin        # :476-488 burn_in_offset
        self._ensure_batch((self.total_steps + total) // self.thin + 1)
        self._batch.run(total)
        self._refresh_stats()
        if self._batch.samples is None:
            return torch.empty((self.num_chains, 0, self.dim), device=self._batch.device, dtype=torch.float32)
        out = self._batch.samples[:, first_row:self._batch.rows_written()]
        return out[0] if self.num_chains == 1 else out

    def run_injected(self, increments, uniforms):
        """Test mode: advance by T steps using the caller's randomness instead of Philox -- increments
        (T, num_chains, dim) AFTER scaling and accept-uniforms (T, num_chains), i.e. exactly what the reference
        pre-generates (`precomputed_increments`, `precomputed_random_vals`, rwm_gpu_optimized.py:490-511).
        Returns the (T, num_chains) uint8 accept decisions."""
        inc = torch.as_tensor(increments)
        T = inc.shape[0]
        self._ensure_batch((self.total_steps + T) // self.thin + 1)
        dec, _ = self._batch.run(T, inj_increments=inc, inj_uniforms=uniforms, want_decisions=True)
        self._refresh_stats()
        return dec

    # ---- accessors (reference: :388-400) ---------------------------------------------------------------
    def get_chain_gpu(self):
        """Chain including the initial state and burn-in."""
        if self._batch is None or self._batch.samples is None:
            return torch.tensor(np.array(self.chain), dtype=self.dtype)
        c = self._batch.samples[:, :self._batch.rows_written()]
        return c[0] if self.num_chains == 1 else c

    def get_log_densities_gpu(self):
        if self._batch is None or self._batch.sample_logp is None:
            return None
        lp = self._batch.sample_logp[:, :self._batch.rows_written()]
        return lp[0] if self.num_chains == 1 else lp

    def esjd_per_chain(self) -> torch.Tensor:
        """Per-chain expected squared jump distance over the post-burn-in steps (float64 tensor)."""
        b = self._batch
        if b is None:
            raise ValueError("The algorithm has not been run yet.")
        post = b.post_burn_in_steps()
        if b.samples is not None and self.thin == 1 and b.rows_written() == b.total_steps + 1:
            if b.rows_written() <= self.burn_in + 1:
                raise ValueError(f"Insufficient post-burn-in samples: chain_index={b.rows_written()}, burn_in={self.burn_in}. "
                                 f"Need at least {self.burn_in + 2} total samples.")
            # the reference's definition on the stored chain (:513-534), via the reduction kernel
            return b.esjd_from_samples(self.burn_in, b.rows_written() - self.burn_in)
        if post < 1:
            raise ValueError("Insufficient post-burn-in samples")
        return b.sq_jump_sum / post

    def expected_squared_jump_distance_gpu(self):
        return float(self.esjd_per_chain().mean().item())

    def get_diagnostic_info(self):
        return {
            'device': str(self.device), 'dtype': str(self.dtype), 'optimization_level': 'SM100A_PERSISTENT_FUSED',
            'use_efficient_rng': self.use_efficient_rng, 'compiled_target': True, 'total_steps': self.total_steps,
            'acceptance_rate': self.acceptance_rate, 'num_chains': self.num_chains,
            'kernel_fusion': 'all steps of all chains in one persistent kernel launch',
            'memory_allocated_mb': torch.cuda.memory_allocated() / 1e6 if torch.cuda.is_available() else 0,
            'random_generation': 'in-kernel Philox4x32-10',
        }

    def performance_comparison_summary(self):
        print(f"{self.name}: persistent fused sm_100a kernel, {self.num_chains} chain(s), "
              f"{self.total_steps} steps, acceptance {self.acceptance_rate:.3f}")


# BASELINE.json's north_star calls the class `RandomWalkMetropolis`; same object under both names.
RandomWalkMetropolis = RandomWalkMH_GPU_Optimized
