"""`ParallelTemperingRWM_GPU_Optimized` -- drop-in for the reference class of the same name
(algorithms/pt_rwm_gpu_optimized.py:101-841), backed by the persistent fused sm_100a kernel with the whole
temperature ladder resident in one CTA.

Same constructor arguments (same order), methods and attributes as the reference.  Appended keyword arguments:
`num_ladders` (independent ladders in one launch), `seed`, `store` ('all' | 'cold' | 'none'), `thin`,
`swap_mode` ('reference' = what the reference really does on an accepted swap: copy k -> j, k unchanged,
pt_rwm_gpu_optimized.py:50-59; 'exchange' = textbook PT), `math_mode`, `proposal_distribution` (Laplace /
UniformRadius ladders, BASELINE config 4 -- an extension, the reference PT takes `var` only), `initial_states`,
`chain_id_base` / `ladder_id_base`, `lanes_per_chain`.
"""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..interfaces import MHAlgorithm, TargetDistribution, TorchTargetDistribution
from ..proposal_distributions import ProposalDistribution, NormalProposal
from ._engine import LadderBatch


class ParallelTemperingRWM_GPU_Optimized(MHAlgorithm):
    """Parallel-tempering RWM for `num_ladders` independent ladders of `len(beta_ladder)` temperatures."""

    def __init__(self, dim: int, var: float,
                 target_dist: TorchTargetDistribution | TargetDistribution = None,
                 symmetric: bool = True,
                 beta_ladder: list = None,
                 iterative_temp_spacing: bool = False,
                 geom_temp_spacing: bool = False,
                 swap_acceptance_rate: float = 0.234,
                 beta_min_iterative: float = 0.01,
                 N_samples_swap_est: int = 3000,
                 iterative_tolerance: float = 0.005,
                 iterative_initial_pn: float = 0.5,
                 iterative_pn_update_power: float = -0.25,
                 iterative_max_pn_steps: int = 100,
                 iterative_pn_clamp_min: float = -10.0,
                 iterative_pn_clamp_max: float = 10.0,
                 iterative_fail_tol_factor: float = 3.0,
                 swap_every: int = 100,
                 burn_in: int = 0,
                 device: str = None,
                 pre_allocate_steps: int = None,
                 dtype: torch.dtype = torch.float32,
                 # ---- appended keyword arguments (not in the reference) ----
                 num_ladders: int = 1,
                 seed: Optional[int] = None,
                 store: Optional[str] = None,
                 thin: int = 1,
                 swap_mode: Optional[str] = None,
                 math_mode: str = "fast",
                 proposal_distribution: ProposalDistribution = None,
                 initial_states=None,
                 chain_id_base: int = 0,
                 lanes_per_chain: int = 0,
                 ladder_id_base: Optional[int] = None):
        if not isinstance(target_dist, TorchTargetDistribution):
            raise TypeError("ParallelTemperingRWM_GPU_Optimized requires a TorchTargetDistribution.")
        self.num_ladders = int(num_ladders)
        if self.num_ladders < 1:
            raise ValueError("num_ladders must be >= 1")
        super().__init__(dim, var, target_dist, symmetric, num_chains=self.num_ladders)
        if device is None:
            self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        else:
            self.device = torch.device(device)
        if dtype != torch.float32:
            raise NotImplementedError("the sm_100a sampling path computes in float32 only")
        self.dtype = dtype
        self.burn_in = max(0, burn_in)
        self.swap_every = swap_every
        self.ideal_swap_acceptance_rate = swap_acceptance_rate
        self.name = "PT_RWM_GPU_ULTRA_FUSED_ITERATIVE_LADDER" if iterative_temp_spacing else "PT_RWM_GPU_ULTRA_FUSED"
        self.use_torch_target = True

        if beta_ladder is not None:
            self.beta_ladder = list(beta_ladder)
        elif iterative_temp_spacing:
            self.beta_ladder = self._construct_iterative_ladder(
                target_swap_acceptance_rate=swap_acceptance_rate, beta_min=beta_min_iterative,
                N_samples_for_swap_estimation=N_samples_swap_est, tolerance=iterative_tolerance,
                initial_pn=iterative_initial_pn, pn_update_power=iterative_pn_update_power,
                max_pn_adjustment_steps=iterative_max_pn_steps,
                pn_clamping_range=(iterative_pn_clamp_min, iterative_pn_clamp_max),
                convergence_failure_tolerance_factor=iterative_fail_tol_factor)
        else:
            self.beta_ladder = self._construct_geometric_ladder()
            if not geom_temp_spacing:
                warnings.warn("No specific ladder construction method chosen. Using geometric spacing as default.")
        self.num_chains = len(self.beta_ladder)
        self.beta_tensor = torch.tensor(self.beta_ladder, dtype=torch.float32)

        if proposal_distribution is None:
            if var is None or var <= 0:
                raise ValueError("var must be positive")
            proposal_distribution = NormalProposal(dim, float(var), 1.0, torch.device("cpu"), torch.float32)
            # per-chain std: sqrt in fp32 of the fp32-rounded var/beta_k (pt_rwm_gpu_optimized.py:453-455)
            self._scales = np.sqrt(np.asarray([np.float32(float(var) / b) for b in self.beta_ladder], dtype=np.float32))
        else:
            self._scales = np.asarray([proposal_distribution.chain_scale(float(b)) for b in self.beta_ladder], dtype=np.float32)
        self.proposal_dist = proposal_distribution

        self.pre_allocate_steps = pre_allocate_steps
        self.thin = max(1, int(thin))
        if store is None:
            store = "all" if (self.num_ladders == 1 or pre_allocate_steps) else "none"
        if store not in ("all", "cold", "none"):
            raise ValueError("store must be 'all', 'cold' or 'none'")
        self.store = store
        # None = the reference's semantics (parity with its recorded numbers), with a one-time warning on the first
        # native-RNG run: that "swap" copies k -> j and does not leave the target invariant (see _warn_copy_swap)
        self._swap_mode_defaulted = swap_mode is None
        self.swap_mode = "reference" if swap_mode is None else swap_mode
        if self.swap_mode not in _lib.SWAP_MODES:
            raise ValueError("swap_mode must be 'reference' or 'exchange'")
        self.math_mode = math_mode
        self.seed = seed
        # global id of this batch's first chain; `ladder_id_base` = the same in units of ladders (the ladder length is only
        # known once the ladder is built, which is why shard helpers pass this one)
        self.chain_id_base = chain_id_base if ladder_id_base is None else int(ladder_id_base) * self.num_chains
        self.lanes_per_chain = lanes_per_chain
        if initial_states is not None:
            x0 = np.asarray(initial_states, dtype=np.float64)
            self._x0_full = np.broadcast_to(x0.reshape(self.num_ladders, -1, dim), (self.num_ladders, self.num_chains, dim)).copy()
        else:
            # all chains of a ladder start from the same point (:481)
            self._x0_full = np.repeat(np.asarray(self._x0, dtype=np.float64)[:, None, :], self.num_chains, axis=1)

        self.num_swap_attempts = 0
        self.num_swap_acceptances = 0
        self.swap_acceptance_rate = 0.0
        self.step_counter = 0
        self.squared_jump_distances = 0.0
        self.pt_esjd = 0.0
        self.swap_acceptance_rates = None
        self.mh_acceptance_rates = None
        self.precomputed_swap_randoms = None
        self.swap_random_index = 0
        self._chain_cache = None
        self._batch: Optional[LadderBatch] = None
        self._make_batch()

    _copy_swap_warned = False

    def _warn_copy_swap(self):
        """The reference's accepted "swap" is a copy: chain j receives chain k's state and k keeps its own
        (pt_rwm_gpu_optimized.py:50-59 on tensor views).  It is the default here because every
        recorded PT number of the reference was produced with it, but the cold chain it yields is biased (RoughCarpet
        mode weights .455/.310/.234 instead of .5/.3/.2).  Say so once per process when the caller
        did not choose a mode."""
        cls = ParallelTemperingRWM_GPU_Optimized
        if self._swap_mode_defaulted and not cls._copy_swap_warned:
            cls._copy_swap_warned = True
            warnings.warn("ParallelTemperingRWM_GPU_Optimized runs with swap_mode='reference': an accepted swap copies the "
                          "hotter chain's state into the colder one (what the reference does) and does not leave the target "
                          "invariant. Pass swap_mode='exchange' for textbook parallel tempering, or swap_mode='reference' "
                          "explicitly to silence this warning.", stacklevel=3)

    # ---- ladder construction ---------------------------------------------------------------------------
    def _construct_geometric_ladder(self):
        """beta = 1, halve while > 0.01, append 0.01 (:245-257)."""
        beta, ladder = 1.0, []
        while beta > 1e-2:
            ladder.append(beta)
            beta = beta * 0.5
        ladder.append(1e-2)
        return ladder

    def _get_typical_samples_at_beta(self, beta_val: float, N_samples: int) -> torch.Tensor:
        if not hasattr(self.target_dist, 'draw_samples_torch'):
            raise NotImplementedError("The target distribution must implement 'draw_samples_torch(n_samples, beta)' "
                                      "for iterative temperature ladder construction.")
        return self.target_dist.draw_samples_torch(N_samples, beta_val)

    # families whose heuristic tempered sampler exists natively (csrc/rwmpt_ladder.cuh, Tempered<...>)
    _NATIVE_LADDER_FAMILIES = (_lib.T_ROUGH_CARPET, _lib.T_THREE_MIXTURE, _lib.T_EVEN_ROSENBROCK, _lib.T_SCALED_MVN, _lib.T_MVN_DIAG)

    def _estimate_swap_probability(self, beta_curr: float, beta_star: float, n: int) -> float:
        """mean(min(1, exp((beta - beta*) (log pi(x*) - log pi(x))))) over heuristic tempered samples (:356-368).

        Native path (`rwmpt_swap_prob_estimate`): ONE kernel draws both sample sets with in-kernel Philox, evaluates both
        log-densities with the sampling kernel's functors and reduces to one fp64 -- nothing but the scalar leaves the
        GPU, and reading it is the only synchronisation of a Robbins-Monro step.  Families without a native tempered
        sampler use the target's `draw_samples_torch` with the CUDA log-density."""
        t = self.target_dist
        if getattr(t, "family_id", None) in self._NATIVE_LADDER_FAMILIES and self.device.type == "cuda" and torch.cuda.is_available():
            dev = _lib.require_cuda(self.device)
            lib = _lib.load()
            if getattr(self, "_ladder_seed", None) is None:
                from ..proposal_distributions.base import draw_seed
                self._ladder_seed, self._ladder_rows = draw_seed(None), 0
            params = t.device_params(dev)
            acc = torch.zeros(1, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.rwmpt_swap_prob_estimate(_lib.target_struct(t.family_id, self.dim, params), float(beta_curr),
                                                        float(beta_star), int(n), self._ladder_seed, self._ladder_rows,
                                                        acc.data_ptr(), _lib.stream_ptr(dev)))
            self._ladder_rows += int(n)
            return float(acc.item()) / int(n)
        xs = self._get_typical_samples_at_beta(beta_star, n)
        xc = self._get_typical_samples_at_beta(beta_curr, n)
        lps, lpc = self.target_dist.log_density(xs), self.target_dist.log_density(xc)
        log_r = (beta_curr - beta_star) * (lps - lpc)
        return torch.mean(torch.exp(torch.clamp_max(log_r, 0.0))).item()

    def _construct_iterative_ladder(self, target_swap_acceptance_rate, beta_min, N_samples_for_swap_estimation, tolerance,
                                    initial_pn, pn_update_power, max_pn_adjustment_steps, pn_clamping_range,
                                    convergence_failure_tolerance_factor) -> list:
        """Iterative ladder of the reference (:283-426): beta* = beta / (1 + exp(p_n)), p_n driven towards the
        target swap rate by a Robbins-Monro recursion; same stopping and acceptance rules."""
        beta_ladder, beta_curr = [1.0], 1.0
        while True:
            if beta_curr <= beta_min + 1e-6:
                break
            pn, n_updates, found = initial_pn, 1, False
            last_star, last_prob, it = -1.0, -1.0, 0
            for it in range(1, max_pn_adjustment_steps + 1):
                cpn = float(np.clip(pn, pn_clamping_range[0], pn_clamping_range[1]))
                if beta_curr < 1e-9:
                    last_star = -1.0
                    break
                beta_star = beta_curr / (1.0 + np.exp(cpn))
                last_star = beta_star
                if beta_star < beta_min:
                    break
                prob = self._estimate_swap_probability(beta_curr, beta_star, N_samples_for_swap_estimation)
                last_prob = prob
                if abs(prob - target_swap_acceptance_rate) <= tolerance:
                    beta_ladder.append(beta_star)
                    beta_curr, found = beta_star, True
                    break
                pn = pn + (n_updates ** pn_update_power) * (prob - target_swap_acceptance_rate)
                n_updates += 1
            if not found:
                if it == max_pn_adjustment_steps and last_star >= beta_min and last_star != -1.0 and \
                        abs(last_prob - target_swap_acceptance_rate) <= tolerance * convergence_failure_tolerance_factor:
                    beta_ladder.append(last_star)
                    beta_curr = last_star
                else:
                    break
        if beta_ladder[-1] > beta_min + 1e-5:
            beta_ladder.append(beta_min)
        return beta_ladder

    # ---- device state -----------------------------------------------------------------------------------
    def _make_batch(self):
        if not torch.cuda.is_available() or self.device.type != "cuda":
            return  # constructed on a CPU-only host: parameters only; any sampling call raises in require_cuda
        K, L = self.num_chains, self.num_ladders
        self._batch = LadderBatch(
            self.target_dist, self.dim, L, K, np.asarray(self.beta_ladder, dtype=np.float32)[None, :],
            self.proposal_dist.family_id, self._scales[None, :], self.proposal_dist.dim_scale(), self._x0_full,
            self.device, burn_in=self.burn_in, swap_every=self.swap_every, swap_mode=self.swap_mode,
            math_mode=self.math_mode, seed=self.seed, chain_id_base=self.chain_id_base,
            lanes_per_chain=self.lanes_per_chain)
        if self.pre_allocate_steps and self.store != "none":
            self._batch.allocate_storage(self.store, self.burn_in + self.pre_allocate_steps + 1, self.thin)   # :459-470
        self._sync_views()

    def _require_batch(self) -> LadderBatch:
        if self._batch is None:
            _lib.require_cuda(self.device)
            self._make_batch()
        return self._batch

    def _sync_views(self):
        b = self._batch
        K, L, d = self.num_chains, self.num_ladders, self.dim
        self.current_states = b.state.view(K, d) if L == 1 else b.state.view(L, K, d)
        self.current_log_densities = b.logp if L == 1 else b.logp.view(L, K)
        self.beta_tensor = self.beta_tensor.to(b.device)
        if b.samples is not None:
            n_per = K if b.store_mode == _lib.STORE_ALL else 1
            self.pre_allocated_chains = b.samples if L == 1 else b.samples.view(L, n_per, b.capacity, d)
            self.pre_allocated_log_densities = b.sample_logp if L == 1 else b.sample_logp.view(L, n_per, b.capacity)
            self.chain_indices = torch.full((K,), b.rows_written(), dtype=torch.long)
        else:
            self.pre_allocated_chains = None
            self.pre_allocated_log_densities = None
            self.chain_indices = None

    @property
    def proposal_covs_chol(self):
        """(K, d, d) Cholesky factors of (var/beta_k) I -- diagonal, kept only as a reference-compatible attribute;
        the kernel multiplies by the scalar std instead of calling bmm (:86-99)."""
        s = torch.as_tensor(self._scales)
        return torch.diag_embed(s[:, None].expand(self.num_chains, self.dim).contiguous())

    def get_name(self):
        return self.name

    def reset(self):
        """Reference `reset` (:525-539) clears counters but not the states; here the chains restart too."""
        self.num_swap_attempts = 0
        self.num_swap_acceptances = 0
        self.swap_acceptance_rate = 0.0
        self.step_counter = 0
        self.squared_jump_distances = 0.0
        self.pt_esjd = 0.0
        self.swap_acceptance_rates = None
        self.mh_acceptance_rates = None
        self._chain_cache = None
        self._batch = None
        self._make_batch()

    def _refresh_stats(self):
        b = self._batch
        K, L = self.num_chains, self.num_ladders
        self.step_counter = b.total_steps
        self._chain_cache = None
        rounds = b.swap_rounds()
        attempts_per_ladder = rounds * (K - 1)
        pair_acc = b.swap_accepts[:, :max(K - 1, 0)].cpu().numpy().astype(np.int64) if K > 1 else np.zeros((L, 0), np.int64)
        self.num_swap_attempts = attempts_per_ladder * L
        self.num_swap_acceptances = int(pair_acc.sum())
        betas = np.asarray(self.beta_ladder, dtype=np.float64)
        dbeta2 = (betas[:-1] - betas[1:]) ** 2
        sq = pair_acc.astype(np.float64) @ dbeta2 if K > 1 else np.zeros(L)
        self.squared_jump_distances = float(sq.sum())
        if attempts_per_ladder > 0:
            self.swap_acceptance_rates = pair_acc.sum(axis=1) / attempts_per_ladder
            self.pair_swap_acceptance_rates = pair_acc / max(rounds, 1)
        if L == 1:
            # the reference refreshes these two only when a swap is accepted (:627-633)
            last = int(b.swap_last_attempt.max().item())
            self.swap_acceptance_rate = self.num_swap_acceptances / last if last > 0 else 0.0
            self.pt_esjd = self.squared_jump_distances / last if last > 0 else 0.0
        elif self.num_swap_attempts > 0:
            self.swap_acceptance_rate = self.num_swap_acceptances / self.num_swap_attempts
            self.pt_esjd = self.squared_jump_distances / self.num_swap_attempts
        post = b.post_burn_in_steps()
        if post > 0:
            self.mh_acceptance_rates = (b.accept_count.to(torch.float64) / post).view(L, K)
        self._sync_views()

    def step(self, step_index: int = None):
        """One Metropolis step of all chains, plus the swap sweep when due (:541-574)."""
        b = self._require_batch()
        self._warn_copy_swap()
        if self.store != "none" and b.samples is None:
            b.allocate_storage(self.store, b.total_steps // self.thin + 2, self.thin)
        elif b.samples is not None and not self.pre_allocate_steps:
            b.grow_storage(max(b.capacity, 2 * (b.total_steps // self.thin + 2)))
        b.run(1)
        self._refresh_stats()

    def generate_samples(self, num_samples: int):
        """Run burn_in + num_samples ladder-steps in ONE kernel launch; returns the cold chain's post-burn-in
        samples (:761-770): (num_samples, dim), or (num_ladders, num_samples, dim) for a batch of ladders, or an
        empty tensor with store='none'."""
        b = self._require_batch()
        self._warn_copy_swap()
        total = self.burn_in + int(num_samples)                         # :707
        if self.store != "none":
            need = (b.total_steps + total) // self.thin + 1
            if b.samples is None:
                b.allocate_storage(self.store, need, self.thin)
            elif not self.pre_allocate_steps:
                b.grow_storage(need)
        b.run(total)
        self._refresh_stats()
        if b.samples is None:
            return torch.empty((self.num_ladders, 0, self.dim), device=b.device, dtype=torch.float32)
        cold = self.get_cold_chain_gpu()
        first = 1 + self.burn_in // self.thin
        return cold[first:] if self.num_ladders == 1 else cold[:, first:]

    def run_injected(self, increments, uniforms, swap_uniforms):
        """Test mode: advance by T ladder-steps with the caller's randomness -- increments (T, L*K, dim) AFTER the
        per-chain scaling (the reference's post-bmm increments, :576-592), accept-uniforms (T, L*K) and one swap
        uniform per (sweep, ladder, pair) (R, L, K-1).  Returns (decisions (T, L*K), swap_decisions (R, L, K-1))."""
        b = self._require_batch()
        inc = torch.as_tensor(increments)
        T = inc.shape[0]
        if self.store != "none":
            need = (b.total_steps + T) // self.thin + 1
            if b.samples is None:
                b.allocate_storage(self.store, need, self.thin)
            elif not self.pre_allocate_steps:
                b.grow_storage(need)
        out = b.run(T, inj_increments=inc, inj_uniforms=uniforms, inj_swap_uniforms=swap_uniforms, want_decisions=True)
        self._refresh_stats()
        return out

    # ---- accessors (:655-692) ------------------------------------------------------------------------------
    def get_all_chains_gpu(self):
        b = self._batch
        if b is None or b.samples is None:
            return []
        rows = b.rows_written()
        if self.num_ladders == 1:
            return [b.samples[i, :rows] for i in range(b.samples.shape[0])]
        n_per = self.num_chains if b.store_mode == _lib.STORE_ALL else 1
        v = b.samples.view(self.num_ladders, n_per, b.capacity, self.dim)
        return [v[:, i, :rows] for i in range(n_per)]

    def get_cold_chain_gpu(self):
        chains = self.get_all_chains_gpu()
        return chains[0] if chains else torch.empty(0, self.dim)

    def _get_cold_chain_cpu(self):
        cold = self.get_cold_chain_gpu()
        if self.num_ladders > 1:
            cold = cold[0]
        return cold.detach().cpu().numpy().tolist()

    @property
    def chain(self):
        if self._batch is None or self._batch.samples is None:
            return [self._x0[0]]
        if self._chain_cache is None:
            self._chain_cache = self._get_cold_chain_cpu()
        return self._chain_cache

    @chain.setter
    def chain(self, value):
        self._chain_cache = value

    def esjd_per_ladder(self) -> torch.Tensor:
        """Cold-chain ESJD per ladder (includes swap moves), float64 tensor of length num_ladders."""
        b = self._require_batch()
        post = b.post_burn_in_steps()
        if b.samples is not None and self.thin == 1 and b.rows_written() == b.total_steps + 1:
            if b.rows_written() <= self.burn_in + 1:
                raise ValueError("Insufficient post-burn-in samples")
            e = b.esjd_from_samples(self.burn_in, b.rows_written() - self.burn_in)        # :772-789
            n_per = self.num_chains if b.store_mode == _lib.STORE_ALL else 1
            return e.view(self.num_ladders, n_per)[:, 0]
        if post < 1:
            raise ValueError("Insufficient post-burn-in samples")
        return b.sq_jump_sum.view(self.num_ladders, self.num_chains)[:, 0] / post

    def expected_squared_jump_distance_gpu(self):
        return float(self.esjd_per_ladder().mean().item())

    # ---- single-phase helpers poked by the reference's debug scripts (tests/debug_pt_performance.py) -----------
    def _generate_all_increments(self):
        b = self._require_batch()
        out = torch.empty((self.num_chains, self.dim), device=b.device, dtype=torch.float32)
        ds = None if b.prop_dim_scale is None else b.prop_dim_scale
        for k in range(self.num_chains):
            _lib.check(b.lib.rwmpt_proposal_sample(b.prop_family, self.dim, float(self._scales[k]), _lib.ptr(ds), 1,
                                                   b.ensure_seed(), (b.total_steps << 8) + k, out[k].data_ptr(),
                                                   _lib.stream_ptr(b.device)))
        return out

    def _compute_log_densities_for_proposals(self, proposals):
        return self.target_dist.log_density(proposals)

    def _attempt_all_swaps(self):
        """One stand-alone sweep over the resident ladders (`rwmpt_pt_swap`)."""
        b = self._require_batch()
        rnd = b.swap_rounds() + getattr(self, "_extra_rounds", 0)
        self._extra_rounds = getattr(self, "_extra_rounds", 0) + 1
        with torch.cuda.device(b.device):
            _lib.check(b.lib.rwmpt_pt_swap(b.state.data_ptr(), b.logp.data_ptr(), b.beta.data_ptr(), self.num_ladders,
                                           self.num_chains, self.dim, b.swap_mode, None, b.ensure_seed(),
                                           b.chain_id_base // self.num_chains, rnd, None, b.swap_accepts.data_ptr(),
                                           _lib.stream_ptr(b.device)))

    def get_diagnostic_info(self):
        return {
            'device': str(self.device), 'dtype': str(self.dtype), 'algorithm': self.name, 'num_chains': self.num_chains,
            'num_ladders': self.num_ladders, 'beta_ladder': self.beta_ladder, 'swap_every': self.swap_every,
            'step_counter': self.step_counter, 'swap_acceptance_rate': self.swap_acceptance_rate, 'pt_esjd': self.pt_esjd,
            'optimization_level': 'SM100A_PERSISTENT_FUSED_LADDER_PER_CTA',
            'memory_allocated_mb': torch.cuda.memory_allocated() / 1e6 if torch.cuda.is_available() else 0,
        }

    def performance_summary(self):
        print(f"{self.name}: {self.num_ladders} ladder(s) x {self.num_chains} temperatures, {self.step_counter} steps, "
              f"swap acceptance {self.swap_acceptance_rate:.3f}, PT-ESJD {self.pt_esjd:.6f}")
