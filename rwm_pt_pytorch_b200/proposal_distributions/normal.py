"""`NormalProposal` (reference: proposal_distributions/normal.py:5-65): increments N(0, (var/beta) I)."""
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .base import ProposalDistribution


class NormalProposal(ProposalDistribution):
    family_id = _lib.P_NORMAL

    def __init__(self, dim: int, base_variance_scalar: float, beta: float, device: torch.device, dtype: torch.dtype,
                 rng_generator: Optional[torch.Generator] = None):
        super().__init__(dim, beta, device, dtype, rng_generator)
        self.name = "Normal"
        if base_variance_scalar <= 0:
            raise ValueError("base_variance_scalar must be positive")
        self.base_variance_scalar = float(base_variance_scalar)
        # normal.py:27-31 -- division in Python float64, cast to fp32, then sqrt in fp32
        effective_variance = base_variance_scalar / self.beta
        self.std_dev = torch.sqrt(torch.tensor(effective_variance, device=self.device, dtype=self.dtype))

    def chain_scale(self, beta: float) -> float:
        return float(np.sqrt(np.float32(self.base_variance_scalar / beta)))

    def get_name(self) -> str:
        return self.name
