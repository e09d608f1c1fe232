"""`LaplaceProposal` (reference: proposal_distributions/laplace.py:5-90): independent Laplace increments with
per-dimension variance var_i / beta, i.e. scale_i = sqrt((var_i / beta) / 2), drawn by inverse CDF."""
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .base import ProposalDistribution


class LaplaceProposal(ProposalDistribution):
    family_id = _lib.P_LAPLACE

    def __init__(self, dim: int, base_variance_vector: torch.Tensor, beta: float, device: torch.device,
                 dtype: torch.dtype, rng_generator: Optional[torch.Generator] = None):
        super().__init__(dim, beta, device, dtype, rng_generator)
        self.name = "Laplace"
        base_variance_vector = torch.as_tensor(base_variance_vector)
        if base_variance_vector.shape != (dim,):
            raise ValueError(f"base_variance_vector must have shape ({dim},), got {base_variance_vector.shape}")
        if not (base_variance_vector > 0).all():
            raise ValueError("All elements of base_variance_vector must be positive")
        self.base_variance_vector = base_variance_vector.detach().to(device="cpu", dtype=torch.float32)
        # laplace.py:29-32 -- fp32 throughout
        effective = base_variance_vector.to(device=self.device, dtype=self.dtype) / self.beta
        self.scale_vector = torch.sqrt(effective / 2.0)

    # kernel: increment_i = prop_scale[chain] * dim_scale[i] * laplace(0, 1)
    def chain_scale(self, beta: float) -> float:
        if beta == self.beta:
            return 1.0
        # a chain at another temperature (PT extension): scale_i(beta') = scale_i(beta) * sqrt(beta / beta')
        return float(np.sqrt(np.float32(self.beta / beta)))

    def dim_scale(self):
        return self.scale_vector.detach().to(dtype=torch.float32)

    def get_name(self) -> str:
        return self.name
