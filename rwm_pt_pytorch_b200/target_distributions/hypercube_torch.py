"""`HypercubeTorch` (reference: target_distributions/hypercube_torch.py:5-113): uniform on [L, R]^d."""
import numpy as np
import torch

from .. import _lib
from ..interfaces.target_torch import TorchTargetDistribution
from ._common import _MoveTensorsMixin, t2n


class HypercubeTorch(_MoveTensorsMixin, TorchTargetDistribution):
    family_id = _lib.T_HYPERCUBE

    def __init__(self, dim, left_boundary=0.0, right_boundary=1.0, device=None):
        super().__init__(dim, device)
        self.name = "HypercubeTorch"
        self.left_boundary = torch.tensor(left_boundary, device=self.device, dtype=torch.float32)
        self.right_boundary = torch.tensor(right_boundary, device=self.device, dtype=torch.float32)
        volume = (right_boundary - left_boundary) ** dim
        self.uniform_density = torch.tensor(1.0 / volume, device=self.device, dtype=torch.float32)
        self.log_uniform_density = torch.log(self.uniform_density)

    def _pack(self):
        return self._header(float(self.left_boundary), float(self.right_boundary), float(self.log_uniform_density))

    def spec(self):
        return dict(family="hypercube", left=t2n(self.left_boundary), right=t2n(self.right_boundary),
                    log_uniform_density=t2n(self.log_uniform_density))

    def get_name(self):
        return self.name

    def draw_sample(self, beta=1.0):
        return np.random.uniform(self.left_boundary.cpu().numpy(), self.right_boundary.cpu().numpy(), self.dim)

    def draw_samples_torch(self, n_samples, beta=1.0):
        s = torch.rand(n_samples, self.dim, device=self.device, dtype=torch.float32)
        return s * (self.right_boundary - self.left_boundary) + self.left_boundary
