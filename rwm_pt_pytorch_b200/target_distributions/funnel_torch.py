"""`NealFunnelTorch` (reference: target_distributions/funnel_torch.py:6-108).
`SuperFunnelTorch` (:111-348, hierarchical logistic regression with per-group data) is out of scope for the
fused kernel and raises NotImplementedError."""
import math

import torch

from .. import _lib
from ..interfaces.target_torch import TorchTargetDistribution
from ._common import _MoveTensorsMixin, t2n


class NealFunnelTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """log p(v, z) = log N(v | mu_v, sigma_v^2) + sum_k log N(z_k | mu_z, e^v)."""
    family_id = _lib.T_NEAL_FUNNEL

    def __init__(self, dim, mu_v=0.0, sigma_v_sq=9.0, mu_z=0.0, device=None):
        super().__init__(dim, device)
        if dim < 1:
            raise ValueError("dim must be at least 1 for Neal's Funnel (v variable).")
        self.mu_v = torch.tensor(mu_v, device=self.device, dtype=torch.float32)
        self.sigma_v_sq = torch.tensor(sigma_v_sq, device=self.device, dtype=torch.float32)
        if self.sigma_v_sq <= 0:
            raise ValueError("sigma_v_sq must be positive.")
        self.mu_z = torch.tensor(mu_z, device=self.device, dtype=torch.float32)
        self.log_sigma_v_sq = torch.log(self.sigma_v_sq)
        self.log_2_pi = torch.tensor(2.0 * math.pi, device=self.device, dtype=torch.float32).log()
        self.D_tensor = torch.tensor(float(self.dim), device=self.device, dtype=torch.float32)
        self.D_minus_1_tensor = torch.tensor(float(max(self.dim - 1, 0)), device=self.device, dtype=torch.float32)

    def _pack(self):
        return self._header(float(self.mu_v), float(self.sigma_v_sq), float(self.mu_z), float(self.log_sigma_v_sq),
                            float(self.log_2_pi), float(self.D_minus_1_tensor))

    def spec(self):
        return dict(family="neal_funnel", mu_v=t2n(self.mu_v), sigma_v_sq=t2n(self.sigma_v_sq), mu_z=t2n(self.mu_z),
                    log_sigma_v_sq=t2n(self.log_sigma_v_sq), log_2pi=t2n(self.log_2_pi), dm1=t2n(self.D_minus_1_tensor))

    def get_name(self):
        return f"NealFunnelTorch_D{self.dim}"

    def draw_sample(self, beta=1.0):
        raise NotImplementedError("NealFunnelTorch.draw_sample is not implemented.")  # as the reference (:87)


class SuperFunnelTorch(TorchTargetDistribution):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("SuperFunnelTorch (data-dependent likelihood) is outside the fused sm_100a "
                                  "sampling path; see DESIGN.md, out of scope.")

    def _pack(self):  # pragma: no cover
        raise NotImplementedError

    def get_name(self):  # pragma: no cover
        return "SuperFunnelTorch"
