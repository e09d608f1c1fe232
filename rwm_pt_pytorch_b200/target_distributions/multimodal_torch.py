"""Multimodal targets (reference: target_distributions/multimodal_torch.py).

`ThreeMixtureDistributionTorch` (:4-348) and `RoughCarpetDistributionTorch` (:351-575): same constructors,
names, attributes and samplers; `log_density` is the CUDA functor (csrc/rwmpt_targets.cuh)."""
import torch

from .. import _lib
from ..interfaces.target_torch import TorchTargetDistribution
from ._common import _MoveTensorsMixin, t2n


def _validate_weights(mode_weights):
    if len(mode_weights) != 3:
        raise ValueError(f"mode_weights must contain exactly 3 weights, got {len(mode_weights)}")
    w = torch.tensor(mode_weights, dtype=torch.float32)
    if not torch.all(w > 0):
        raise ValueError("All mode_weights must be positive")
    if not torch.allclose(torch.sum(w), torch.tensor(1.0), rtol=1e-6):
        raise ValueError(f"mode_weights must sum to 1.0, got sum = {torch.sum(w).item()}")


class ThreeMixtureDistributionTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """p(x) = sum_k w_k N(x | mu_k, I)  (optionally in coordinates scaled by s, times prod s)."""
    family_id = _lib.T_THREE_MIXTURE

    def __init__(self, dim, scaling=False, device=None, mode_centers=None, mode_weights=None):
        super().__init__(dim, device)
        if mode_centers is None:
            mode_centers = [[-5.0] + [0.0] * (dim - 1), [0.0] * dim, [5.0] + [0.0] * (dim - 1)]
        if mode_weights is None:
            mode_weights = [1 / 3, 1 / 3, 1 / 3]
        if len(mode_centers) != 3:
            raise ValueError(f"mode_centers must contain exactly 3 modes, got {len(mode_centers)}")
        for i, center in enumerate(mode_centers):
            if len(center) != dim:
                raise ValueError(f"Mode {i} has dimension {len(center)}, expected {dim}")
        _validate_weights(mode_weights)
        self.means = torch.tensor(mode_centers, device=self.device, dtype=torch.float32)
        self.mixing_weights = torch.tensor(mode_weights, device=self.device, dtype=torch.float32)
        self.log_mixing_weights = torch.log(self.mixing_weights)
        _log_2pi = torch.log(torch.tensor(2.0 * torch.pi, device=self.device, dtype=torch.float32))
        self.scaling_arg_from_constructor = bool(scaling)
        self.cov_dets = torch.ones(3, device=self.device, dtype=torch.float32)
        self.log_norm_consts = -0.5 * (self.dim * _log_2pi + torch.log(self.cov_dets))   # :78 / :98-102
        if scaling:
            self.scaling_factors = torch.rand(dim, device=self.device, dtype=torch.float32) * (1.98 - 0.02) + 0.02
            self.log_jacobian = torch.sum(torch.log(self.scaling_factors))
            self.base_log_norm_const_for_scaled = -0.5 * self.dim * _log_2pi
        default_centers = torch.tensor([[-5.0] + [0.0] * (dim - 1), [0.0] * dim, [5.0] + [0.0] * (dim - 1)])
        is_default = (torch.allclose(torch.tensor(mode_centers, dtype=torch.float32), default_centers, rtol=1e-6)
                      and torch.allclose(torch.tensor(mode_weights, dtype=torch.float32),
                                         torch.tensor([1 / 3, 1 / 3, 1 / 3]), rtol=1e-6))
        self.name = "ThreeMixtureTorch" + ("" if is_default else "Custom") + ("Scaled" if scaling else "")

    def _pack(self):
        if self.scaling_arg_from_constructor:
            c1 = [float(self.base_log_norm_const_for_scaled)] * 3
            head = self._header(*self.log_mixing_weights.tolist(), *c1, 1.0, float(self.log_jacobian))
            return torch.cat([head, self.means.cpu().reshape(-1), self.scaling_factors.cpu()])
        head = self._header(*self.log_mixing_weights.tolist(), *self.log_norm_consts.tolist(), 0.0, 0.0)
        return torch.cat([head, self.means.cpu().reshape(-1)])

    def spec(self):
        s = dict(family="three_mixture", means=t2n(self.means), log_weights=t2n(self.log_mixing_weights))
        if self.scaling_arg_from_constructor:
            s.update(scaling=t2n(self.scaling_factors), log_jacobian=t2n(self.log_jacobian),
                     c1=t2n(self.base_log_norm_const_for_scaled).repeat(3))
        else:
            s["c1"] = t2n(self.log_norm_consts)
        return s

    def get_name(self):
        return self.name

    def draw_samples_torch(self, n_samples, beta=1.0):
        """Heuristic tempered sampler used by the iterative ladder (multimodal_torch.py:290-333)."""
        idx = torch.multinomial(self.mixing_weights, n_samples, replacement=True)
        beta_t = torch.tensor(beta, device=self.device, dtype=torch.float32)
        y = self.means[idx] + torch.randn(n_samples, self.dim, device=self.device, dtype=torch.float32) / torch.sqrt(beta_t)
        if self.scaling_arg_from_constructor:
            y = y / self.scaling_factors.unsqueeze(0)
        return y

    def draw_sample(self, beta=1.0):
        return self.draw_samples_torch(1, beta)[0].cpu().numpy()


class RoughCarpetDistributionTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """Product over coordinates of a 1-D three-mode Gaussian mixture."""
    family_id = _lib.T_ROUGH_CARPET

    def __init__(self, dim, scaling=False, device=None, mode_centers=None, mode_weights=None):
        super().__init__(dim, device)
        if mode_centers is None:
            mode_centers = [-5.0, 0.0, 5.0]
        if mode_weights is None:
            mode_weights = [0.5, 0.3, 0.2]
        if len(mode_centers) != 3:
            raise ValueError(f"mode_centers must contain exactly 3 modes, got {len(mode_centers)}")
        for i, center in enumerate(mode_centers):
            if not isinstance(center, (int, float)):
                raise ValueError(f"Mode center {i} must be a scalar, got {type(center)}")
        _validate_weights(mode_weights)
        is_default = (torch.allclose(torch.tensor(mode_centers, dtype=torch.float32), torch.tensor([-5.0, 0.0, 5.0]), rtol=1e-6)
                      and torch.allclose(torch.tensor(mode_weights, dtype=torch.float32), torch.tensor([0.5, 0.3, 0.2]), rtol=1e-6))
        self.name = "RoughCarpetTorch" + ("" if is_default else "Custom") + ("Scaled" if scaling else "")
        self.modes = torch.tensor(mode_centers, device=self.device, dtype=torch.float32)
        self.weights = torch.tensor(mode_weights, device=self.device, dtype=torch.float32)
        self.log_weights = torch.log(self.weights)
        self.log_sqrt_2pi = torch.log(torch.sqrt(torch.tensor(2.0 * torch.pi, device=self.device, dtype=torch.float32)))
        if scaling:
            self.scaling_factors = torch.rand(dim, device=self.device, dtype=torch.float32) * (1.98 - 0.02) + 0.02

    def _pack(self):
        has_s = hasattr(self, 'scaling_factors')
        jac = float(torch.sum(torch.log(self.scaling_factors.cpu()))) if has_s else 0.0   # :486
        head = self._header(*self.modes.tolist(), *self.log_weights.tolist(), float(self.log_sqrt_2pi),
                            1.0 if has_s else 0.0, jac)
        return torch.cat([head, self.scaling_factors.cpu()]) if has_s else head

    def spec(self):
        s = dict(family="rough_carpet", modes=t2n(self.modes), log_weights=t2n(self.log_weights),
                 log_sqrt_2pi=t2n(self.log_sqrt_2pi))
        if hasattr(self, 'scaling_factors'):
            s["scaling"] = t2n(self.scaling_factors)
        return s

    def get_name(self):
        return self.name

    def draw_samples_torch(self, n_samples, beta=1.0):
        """multimodal_torch.py:532-565."""
        idx = torch.multinomial(self.weights, n_samples * self.dim, replacement=True).view(n_samples, self.dim)
        beta_t = torch.tensor(beta, device=self.device, dtype=torch.float32)
        samples = self.modes[idx] + torch.randn(n_samples, self.dim, device=self.device, dtype=torch.float32) / torch.sqrt(beta_t)
        if hasattr(self, 'scaling_factors'):
            samples = samples / self.scaling_factors.unsqueeze(0)
        return samples

    def draw_sample(self, beta=1.0):
        return self.draw_samples_torch(1, beta)[0].cpu().numpy()
