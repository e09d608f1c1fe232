"""Rosenbrock targets of Pagani et al. (2022) (reference: target_distributions/rosenbrock_torch.py):
`FullRosenbrockTorch` (:13-130), `EvenRosenbrockTorch` (:133-261), `HybridRosenbrockTorch` (:264-410)."""
from typing import Union

import numpy as np
import torch

from .. import _lib
from ..interfaces.target_torch import TorchTargetDistribution
from ._common import _MoveTensorsMixin, t2n

DEFAULT_A_COEFF = 1.0 / 20.0
DEFAULT_B_COEFF = 100.0 / 20.0
DEFAULT_MU = 1.0


def _mu_vector(mu, n, device):
    if isinstance(mu, (int, float)):
        return torch.full((n,), mu, device=device, dtype=torch.float32)
    if isinstance(mu, torch.Tensor):
        if mu.ndim == 0:
            return torch.full((n,), mu.item(), device=device, dtype=torch.float32)
        if mu.shape == (n,):
            return mu.to(device=device, dtype=torch.float32)
        raise ValueError(f"mu tensor must be scalar or have shape ({n},)")
    raise TypeError("mu must be float, int, or torch.Tensor")


class FullRosenbrockTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """log p(x) = -sum_{i<d-1} [b (x_{i+1} - x_i^2)^2 + a (x_i - mu_i)^2]."""
    family_id = _lib.T_FULL_ROSENBROCK

    def __init__(self, dim: int, a_coeff: float = DEFAULT_A_COEFF, b_coeff: float = DEFAULT_B_COEFF,
                 mu: Union[float, torch.Tensor] = DEFAULT_MU, device: str = None):
        if dim < 2:
            raise ValueError("Dimension for FullRosenbrockTorch must be at least 2.")
        super().__init__(dim, device)
        self.a_coeff = torch.tensor(a_coeff, device=self.device, dtype=torch.float32)
        self.b_coeff = torch.tensor(b_coeff, device=self.device, dtype=torch.float32)
        self.mu = _mu_vector(mu, dim - 1, self.device)
        self._name = "FullRosenbrockTorch"

    def _pack(self):
        return torch.cat([self._header(float(self.a_coeff), float(self.b_coeff)), self.mu.cpu()])

    def spec(self):
        return dict(family="full_rosenbrock", a=t2n(self.a_coeff), b=t2n(self.b_coeff), mu=t2n(self.mu))

    def get_name(self) -> str:
        return self._name

    def draw_samples_torch(self, n_samples: int, beta: float = 1.0) -> torch.Tensor:
        raise NotImplementedError("Draw samples for FullRosenbrockTorch is not implemented yet.")  # as the reference (:101)


class EvenRosenbrockTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """log p(x) = -sum_j [a (x_{2j} - mu_j)^2 + b (x_{2j+1} - x_{2j}^2)^2], d even."""
    family_id = _lib.T_EVEN_ROSENBROCK

    def __init__(self, dim: int, a_coeff: float = DEFAULT_A_COEFF, b_coeff: float = DEFAULT_B_COEFF,
                 mu: Union[float, torch.Tensor] = DEFAULT_MU, device: str = None):
        if dim < 2 or dim % 2 != 0:
            raise ValueError("Dimension for EvenRosenbrockTorch must be at least 2 and even.")
        super().__init__(dim, device)
        self.a_coeff = torch.tensor(a_coeff, device=self.device, dtype=torch.float32)
        self.b_coeff = torch.tensor(b_coeff, device=self.device, dtype=torch.float32)
        self.mu = _mu_vector(mu, dim // 2, self.device)
        self._name = "EvenRosenbrockTorch"

    def _pack(self):
        return torch.cat([self._header(float(self.a_coeff), float(self.b_coeff)), self.mu.cpu()])

    def spec(self):
        return dict(family="even_rosenbrock", a=t2n(self.a_coeff), b=t2n(self.b_coeff), mu=t2n(self.mu))

    def get_name(self) -> str:
        return self._name

    def draw_samples_torch(self, n_samples: int, beta: float = 1.0) -> torch.Tensor:
        """Conditional-Gaussian heuristic (rosenbrock_torch.py:224-248)."""
        samples = torch.zeros(n_samples, self.dim, device=self.device, dtype=torch.float32)
        eff_a, eff_b = self.a_coeff * beta, self.b_coeff * beta
        pairs = self.dim // 2
        var_odd = 1.0 / (2 * eff_a) if eff_a > 0 else 1.0
        odd = self.mu.expand(n_samples, pairs) + torch.randn(n_samples, pairs, device=self.device) * torch.sqrt(var_odd)
        var_even = 1.0 / (2 * eff_b) if eff_b > 0 else 1.0
        even = odd ** 2 + torch.randn(n_samples, pairs, device=self.device) * torch.sqrt(var_even)
        samples[:, 0::2] = odd
        samples[:, 1::2] = even
        return samples

    def draw_sample(self, beta: float = 1.0) -> np.ndarray:
        return self.draw_samples_torch(1, beta)[0].cpu().numpy()


class HybridRosenbrockTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """log p(x) = -a (x_0 - mu)^2 - b sum_j (x_{j,2} - x_0^2)^2 - b sum_j sum_i (x_{j,i} - x_{j,i-1}^2)^2,
    d = 1 + n2 (n1 - 1)."""
    family_id = _lib.T_HYBRID_ROSENBROCK

    def __init__(self, n1: int, n2: int, a_coeff: float = DEFAULT_A_COEFF, b_coeff: float = DEFAULT_B_COEFF,
                 mu: float = DEFAULT_MU, device: str = None):
        if n1 < 2:
            raise ValueError("n1 (block length parameter) must be at least 2.")
        if n2 < 1:
            raise ValueError("n2 (number of blocks) must be at least 1.")
        super().__init__(1 + n2 * (n1 - 1), device)
        self.n1, self.n2 = n1, n2
        self.a_coeff = torch.tensor(a_coeff, device=self.device, dtype=torch.float32)
        self.b_coeff = torch.tensor(b_coeff, device=self.device, dtype=torch.float32)
        self.mu = torch.tensor(mu, device=self.device, dtype=torch.float32)
        self._name = f"HybridRosenbrockTorch(n1={n1}, n2={n2}, a={a_coeff:.2f}, b={b_coeff:.2f}, mu={mu:.2f})"

    def _pack(self):
        return self._header(float(self.a_coeff), float(self.b_coeff), float(self.mu), self.n1, self.n2)

    def spec(self):
        return dict(family="hybrid_rosenbrock", a=t2n(self.a_coeff), b=t2n(self.b_coeff), mu=t2n(self.mu),
                    n1=self.n1, n2=self.n2)

    def get_name(self) -> str:
        return self._name
