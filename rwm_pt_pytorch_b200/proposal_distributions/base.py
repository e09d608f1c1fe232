"""`ProposalDistribution` plugin base (reference: proposal_distributions/base.py:7-57).

Same constructor and attributes as the reference.  A proposal here is a parameter carrier for the fused
sampling kernel (`kernel_spec`) and `sample(n)` runs the stand-alone CUDA sampler `rwmpt_proposal_sample`
(in-kernel Philox4x32-10; nothing is drawn on the host)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch

from .. import _lib


def draw_seed(rng_generator: Optional[torch.Generator] = None) -> int:
    """A 63-bit Philox key taken from the given torch generator (or the global CPU one), so that
    `torch.manual_seed` controls the stream like it does for the reference's PT sampler."""
    if rng_generator is not None:
        dev = rng_generator.device
        return int(torch.randint(0, 2 ** 62, (1,), generator=rng_generator, device=dev, dtype=torch.int64).item())
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class ProposalDistribution(ABC):
    family_id: int = -1

    def __init__(self, dim: int, beta: float, device: torch.device, dtype: torch.dtype,
                 rng_generator: Optional[torch.Generator] = None):
        self.dim = dim
        self.beta = beta
        self.device = device
        self.dtype = dtype
        self.rng_generator = rng_generator
        self._seed = None
        self._rows_drawn = 0

    # ---- what the fused kernel needs ------------------------------------------------------------
    @abstractmethod
    def chain_scale(self, beta: float) -> float:
        """Per-chain scalar of include/rwmpt.h `prop_scale` for a chain at inverse temperature `beta`."""

    def dim_scale(self) -> Optional[torch.Tensor]:
        """Per-dimension factor (`prop_dim_scale`) or None."""
        return None

    @abstractmethod
    def get_name(self) -> str:
        pass

    # ---- reference API -----------------------------------------------------------------------------
    def sample(self, n_samples: int) -> torch.Tensor:
        """(n_samples, dim) proposal increments, generated on the GPU."""
        dev = _lib.require_cuda(self.device if torch.device(self.device).type == "cuda" else "cuda")
        lib = _lib.load()
        if self._seed is None:
            self._seed = draw_seed(self.rng_generator)
        out = torch.empty((n_samples, self.dim), device=dev, dtype=torch.float32)
        ds = self.dim_scale()
        ds = None if ds is None else ds.to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.rwmpt_proposal_sample(self.family_id, self.dim, float(self._sample_scale()), _lib.ptr(ds),
                                                 n_samples, self._seed, self._rows_drawn, out.data_ptr(),
                                                 _lib.stream_ptr(dev)))
        self._rows_drawn += n_samples
        return out.to(self.dtype)

    def _sample_scale(self) -> float:
        return self.chain_scale(self.beta)

    def sample_into(self, n_samples: int, output_tensor: torch.Tensor) -> None:
        output_tensor.copy_(self.sample(n_samples))
