"""`UniformRadiusProposal` (reference: proposal_distributions/uniform.py:6-100): increments uniform in the
d-ball of radius r / sqrt(beta), as direction z/||z|| times radius R u^(1/d)."""
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .base import ProposalDistribution


class UniformRadiusProposal(ProposalDistribution):
    family_id = _lib.P_UNIFORM_RADIUS

    def __init__(self, dim: int, base_radius: float, beta: float, device: torch.device, dtype: torch.dtype,
                 rng_generator: Optional[torch.Generator] = None):
        super().__init__(dim, beta, device, dtype, rng_generator)
        self.name = "UniformRadius"
        if base_radius <= 0:
            raise ValueError("base_radius must be positive")
        self.base_radius = float(base_radius)
        # uniform.py:28-32 -- r / sqrt(fp32(beta))
        self.effective_radius = base_radius / torch.sqrt(torch.tensor(self.beta, device=self.device, dtype=self.dtype))
        self.inv_dim = 1.0 / self.dim

    def chain_scale(self, beta: float) -> float:
        return float(np.float32(self.base_radius) / np.sqrt(np.float32(beta)))

    def get_name(self) -> str:
        return self.name
