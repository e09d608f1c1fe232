"""Shared host-side engine of the two sampler facades: owns the device tensors of a batch of ladders
(RWM = ladders of one temperature) and drives `rwmpt_rwm_run` / `rwmpt_pt_run` through ctypes.

Nothing here computes on the host: states, log-densities, accumulators and retained samples live in HBM as
torch tensors (torch = device memory + streams only) and every sampling step runs in librwmpt.so."""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..proposal_distributions.base import draw_seed


class LadderBatch:
    """n_ladders x n_temps chains of dimension dim resident on one CUDA device."""

    def __init__(self, target, dim: int, n_ladders: int, n_temps: int, betas, prop_family: int, prop_scale,
                 prop_dim_scale, x0, device, burn_in: int = 0, swap_every: int = 1, swap_mode: str = "reference",
                 math_mode: str = "fast", seed: Optional[int] = None, chain_id_base: int = 0,
                 lanes_per_chain: int = 0, rng_generator=None):
        self.device = _lib.require_cuda(device)
        self.lib = _lib.load()
        self.target = target
        self.dim, self.L, self.K = int(dim), int(n_ladders), int(n_temps)
        self.n_chains = self.L * self.K
        self.burn_in = int(burn_in)
        self.swap_every = int(swap_every)
        self.swap_mode = _lib.SWAP_MODES[swap_mode]
        self.math_mode = _lib.MATH_MODES[math_mode]
        self.prop_family = int(prop_family)
        self.seed = seed
        self.rng_generator = rng_generator
        self.chain_id_base = int(chain_id_base)
        self.lanes_per_chain = int(lanes_per_chain)
        # how units of work are placed on the SMs (include/rwmpt.h RWMPT_SCHEDULE_*: 0 auto, 1 plain, 2 balanced);
        # results never depend on it.  RWMPT_SCHEDULE overrides the default for A/B measurements.
        self.schedule = int(os.environ.get("RWMPT_SCHEDULE", "0"))
        dev = self.device
        f32 = dict(device=dev, dtype=torch.float32)
        self.beta = torch.as_tensor(np.broadcast_to(np.asarray(betas, dtype=np.float32), (self.L, self.K)).copy()).to(dev).reshape(-1).contiguous()
        self.prop_scale = torch.as_tensor(np.broadcast_to(np.asarray(prop_scale, dtype=np.float32), (self.L, self.K)).copy()).to(dev).reshape(-1).contiguous()
        self.prop_dim_scale = None if prop_dim_scale is None else torch.as_tensor(prop_dim_scale).detach().to(**f32).contiguous()
        x0 = torch.as_tensor(np.asarray(x0, dtype=np.float32))
        self.state = x0.to(dev).reshape(self.n_chains, self.dim).contiguous().clone()
        self.params = target.device_params(dev)
        self.logp = torch.empty(self.n_chains, **f32)
        self._eval_logp()
        self.accept_count = torch.zeros(self.n_chains, device=dev, dtype=torch.int64)
        self.sq_jump_sum = torch.zeros(self.n_chains, device=dev, dtype=torch.float64)
        self.swap_accepts = torch.zeros((self.L, max(self.K - 1, 1)), device=dev, dtype=torch.int64)
        self.swap_last_attempt = torch.zeros(self.n_chains, device=dev, dtype=torch.int64)
        self.total_steps = 0
        self.samples = None
        self.sample_logp = None
        self.store_mode = _lib.STORE_NONE
        self.thin = 1
        self.capacity = 0
        self.store_origin = 0     # global step the sample buffer's row 0 corresponds to (see rewind_storage)

    # ------------------------------------------------------------------------------------------------
    def _target_struct(self):
        return _lib.target_struct(self.target.family_id, self.dim, self.params)

    def _eval_logp(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rwmpt_log_density(self._target_struct(), self.state.data_ptr(), self.n_chains,
                                                  self.logp.data_ptr(), self.math_mode, _lib.stream_ptr(self.device)))

    def ensure_seed(self):
        if self.seed is None:
            self.seed = draw_seed(self.rng_generator)
        return self.seed

    def allocate_storage(self, store: str, capacity_rows: int, thin: int = 1, with_logp: bool = True):
        """(stored_chains, capacity_rows, dim) sample buffer; row 0 = initial state (as the reference's
        pre-allocated chains, rwm_gpu_optimized.py:226-238, pt_rwm_gpu_optimized.py:459-470)."""
        self.store_mode = _lib.STORE_MODES[store]
        self.thin = int(thin)
        if self.store_mode == _lib.STORE_NONE:
            self.samples = self.sample_logp = None
            self.capacity = 0
            return
        n_stored = self.n_chains if self.store_mode == _lib.STORE_ALL else self.L
        self.capacity = int(capacity_rows)
        self.samples = torch.zeros((n_stored, self.capacity, self.dim), device=self.device, dtype=torch.float32)
        self.sample_logp = torch.zeros((n_stored, self.capacity), device=self.device, dtype=torch.float32) if with_logp else None
        src = self.state if self.store_mode == _lib.STORE_ALL else self.state.view(self.L, self.K, self.dim)[:, 0]
        self.samples[:, 0] = src
        if with_logp:
            self.sample_logp[:, 0] = self.logp if self.store_mode == _lib.STORE_ALL else self.logp.view(self.L, self.K)[:, 0]

    def grow_storage(self, capacity_rows: int):
        if self.samples is None or capacity_rows <= self.capacity:
            return
        new = torch.zeros((self.samples.shape[0], capacity_rows, self.dim), device=self.device, dtype=torch.float32)
        new[:, :self.capacity] = self.samples
        self.samples = new
        if self.sample_logp is not None:
            nl = torch.zeros((self.samples.shape[0], capacity_rows), device=self.device, dtype=torch.float32)
            nl[:, :self.capacity] = self.sample_logp
            self.sample_logp = nl
        self.capacity = int(capacity_rows)

    def rows_written(self) -> int:
        """Rows of the sample buffer that hold data (initial state + retained steps), capped at capacity."""
        if self.samples is None:
            return 0
        return min(1 + (self.total_steps - self.store_origin) // self.thin, self.capacity)

    def rewind_storage(self):
        """Start filling the sample buffer from row 0 again: row 0 <- the current state, later rows <- the steps that
        follow (`store_start` of include/rwmpt.h).  Lets a long run stream trajectories through a fixed-size buffer, one
        launch per buffer-full (the benchmark's stored-trajectory workloads re-use one buffer for every timed launch)."""
        if self.samples is None:
            return
        self.store_origin = self.total_steps
        src = self.state if self.store_mode == _lib.STORE_ALL else self.state.view(self.L, self.K, self.dim)[:, 0]
        self.samples[:, 0] = src
        if self.sample_logp is not None:
            self.sample_logp[:, 0] = self.logp if self.store_mode == _lib.STORE_ALL else self.logp.view(self.L, self.K)[:, 0]

    # ------------------------------------------------------------------------------------------------
    def run(self, n_steps: int, inj_increments=None, inj_uniforms=None, inj_swap_uniforms=None,
            want_decisions: bool = False):
        """Advance every chain by n_steps (one kernel launch).  Optional injected randomness (test mode):
        inj_increments (n_steps, n_chains, dim), inj_uniforms (n_steps, n_chains),
        inj_swap_uniforms (n_rounds, n_ladders, n_temps-1).  Returns (decisions, swap_decisions) or None."""
        n_steps = int(n_steps)
        dev = self.device
        a = _lib.RunArgs()
        a.target = self._target_struct()
        a.proposal_family = self.prop_family
        a.n_temps = self.K
        a.prop_scale = self.prop_scale.data_ptr()
        a.prop_dim_scale = _lib.ptr(self.prop_dim_scale)
        a.beta = self.beta.data_ptr()
        a.n_ladders, a.n_steps, a.burn_in, a.step_offset = self.L, n_steps, self.burn_in, self.total_steps
        a.swap_every, a.swap_mode = max(self.swap_every, 1), self.swap_mode
        a.state, a.logp = self.state.data_ptr(), self.logp.data_ptr()
        a.seed = self.ensure_seed() if inj_increments is None else 0
        a.chain_id_base = self.chain_id_base
        a.store_mode, a.math_mode = self.store_mode, self.math_mode
        a.store_start, a.thin = self.store_origin, self.thin
        if self.samples is not None:
            # row 0 of the buffer is the initial state: hand the kernel a pointer to row 1
            a.samples = self.samples.data_ptr() + 4 * self.dim
            a.sample_logp = None if self.sample_logp is None else self.sample_logp.data_ptr() + 4
            a.sample_stride = self.capacity
            a.sample_rows = self.capacity - 1
        a.accept_count = self.accept_count.data_ptr()
        a.sq_jump_sum = self.sq_jump_sum.data_ptr()
        a.swap_accepts = self.swap_accepts.data_ptr()
        a.swap_last_attempt = self.swap_last_attempt.data_ptr()
        a.lanes_per_chain = self.lanes_per_chain
        a.schedule = self.schedule
        keep = []
        if inj_increments is not None:
            inc = torch.as_tensor(inj_increments).to(device=dev, dtype=torch.float32).reshape(n_steps, self.n_chains, self.dim).contiguous()
            uu = torch.as_tensor(inj_uniforms).to(device=dev, dtype=torch.float32).reshape(n_steps, self.n_chains).contiguous()
            a.inj_increments, a.inj_uniforms = inc.data_ptr(), uu.data_ptr()
            keep += [inc, uu]
        n_rounds = int(self.lib.rwmpt_count_swap_rounds(self.total_steps, n_steps, self.burn_in, max(self.swap_every, 1))) if self.K > 1 else 0
        if inj_swap_uniforms is not None and self.K > 1:
            su = torch.as_tensor(inj_swap_uniforms).to(device=dev, dtype=torch.float32).contiguous()
            if su.numel() < n_rounds * self.L * (self.K - 1):
                raise ValueError(f"inj_swap_uniforms holds {su.numel()} values, the run needs {n_rounds * self.L * (self.K - 1)}")
            a.inj_swap_uniforms = su.data_ptr()
            keep.append(su)
        dec = sdec = None
        if want_decisions:
            dec = torch.zeros((n_steps, self.n_chains), device=dev, dtype=torch.uint8)
            a.decisions = dec.data_ptr()
            if self.K > 1:
                sdec = torch.zeros((max(n_rounds, 1), self.L, self.K - 1), device=dev, dtype=torch.uint8)
                a.swap_decisions = sdec.data_ptr()
        fn = self.lib.rwmpt_rwm_run if self.K == 1 else self.lib.rwmpt_pt_run
        with torch.cuda.device(dev):
            _lib.check(fn(C.byref(a), _lib.stream_ptr(dev)))
        self.total_steps += n_steps
        self._keepalive = keep  # until the stream has consumed them
        if want_decisions:
            return dec, (None if sdec is None else sdec[:n_rounds])
        return None

    def geometry(self):
        """(elements_per_lane, lanes_per_chain) the fast-math launch of this batch uses (`rwmpt_pick_geometry`)."""
        a = _lib.RunArgs()
        a.target = self._target_struct()
        a.proposal_family, a.n_temps, a.n_ladders = self.prop_family, self.K, self.L
        a.math_mode, a.lanes_per_chain = self.math_mode, self.lanes_per_chain
        w, e = C.c_int32(), C.c_int32()
        _lib.check(self.lib.rwmpt_pick_geometry(C.byref(a), C.byref(w), C.byref(e)))
        return e.value, w.value

    # ---- statistics ------------------------------------------------------------------------------------
    def post_burn_in_steps(self) -> int:
        return max(self.total_steps - self.burn_in, 0)

    def swap_rounds(self) -> int:
        if self.K == 1:
            return 0
        return int(self.lib.rwmpt_count_swap_rounds(0, self.total_steps, self.burn_in, max(self.swap_every, 1)))

    def esjd_from_samples(self, first_row: int, n_rows: int) -> torch.Tensor:
        """Per stored chain: mean squared jump between consecutive retained rows [first_row, first_row+n_rows)
        (`rwmpt_esjd_reduce`, the reduction kernel behind expected_squared_jump_distance_gpu)."""
        out = torch.zeros(self.samples.shape[0], device=self.device, dtype=torch.float64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rwmpt_esjd_reduce(self.samples.data_ptr(), self.samples.shape[0], self.capacity,
                                                  first_row, n_rows, self.dim, out.data_ptr(), None,
                                                  _lib.stream_ptr(self.device)))
        return out
