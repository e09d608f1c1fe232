"""IID product targets (reference: target_distributions/iid_product_torch.py): `IIDGammaTorch` (:5-132) and
`IIDBetaTorch` (:135-273); -inf outside the open support."""
import numpy as np
import torch

from .. import _lib
from ..interfaces.target_torch import TorchTargetDistribution
from ._common import _MoveTensorsMixin, t2n


class IIDGammaTorch(_MoveTensorsMixin, TorchTargetDistribution):
    family_id = _lib.T_IID_GAMMA

    def __init__(self, dim, shape=2.0, scale=3.0, device=None):
        super().__init__(dim, device)
        self.name = "IIDGammaTorch"
        self.shape = torch.tensor(shape, device=self.device, dtype=torch.float32)
        self.scale = torch.tensor(scale, device=self.device, dtype=torch.float32)
        self.log_gamma_shape = torch.lgamma(self.shape)
        self.log_norm_const_1d = self.log_gamma_shape + self.shape * torch.log(self.scale)
        self.log_norm_const = dim * self.log_norm_const_1d

    def _pack(self):
        return self._header(float(self.shape), float(self.scale), float(self.log_norm_const))

    def spec(self):
        return dict(family="iid_gamma", shape=t2n(self.shape), scale=t2n(self.scale), log_norm_const=t2n(self.log_norm_const))

    def get_name(self):
        return self.name

    def draw_sample(self, beta=1.0):
        return np.random.gamma(self.shape.cpu().numpy() * beta, self.scale.cpu().numpy(), self.dim)

    def draw_samples_torch(self, n_samples, beta=1.0):
        dist = torch.distributions.Gamma(self.shape * beta, 1.0 / self.scale)
        return dist.sample((n_samples, self.dim)).to(self.device)


class IIDBetaTorch(_MoveTensorsMixin, TorchTargetDistribution):
    family_id = _lib.T_IID_BETA

    def __init__(self, dim, alpha=2.0, beta=3.0, device=None):
        super().__init__(dim, device)
        self.name = "IIDBetaTorch"
        self.alpha = torch.tensor(alpha, device=self.device, dtype=torch.float32)
        self.beta = torch.tensor(beta, device=self.device, dtype=torch.float32)
        self.log_gamma_alpha = torch.lgamma(self.alpha)
        self.log_gamma_beta = torch.lgamma(self.beta)
        self.log_gamma_alpha_beta = torch.lgamma(self.alpha + self.beta)
        self.log_norm_const_1d = self.log_gamma_alpha_beta - self.log_gamma_alpha - self.log_gamma_beta
        self.log_norm_const = dim * self.log_norm_const_1d

    def _pack(self):
        return self._header(float(self.alpha), float(self.beta), float(self.log_norm_const))

    def spec(self):
        return dict(family="iid_beta", alpha=t2n(self.alpha), beta=t2n(self.beta), log_norm_const=t2n(self.log_norm_const))

    def get_name(self):
        return self.name

    def draw_sample(self, beta_temp=1.0):
        return np.random.beta(self.alpha.cpu().numpy() * beta_temp, self.beta.cpu().numpy() * beta_temp, self.dim)

    def draw_samples_torch(self, n_samples, beta_temp=1.0):
        dist = torch.distributions.Beta(self.alpha * beta_temp, self.beta * beta_temp)
        return dist.sample((n_samples, self.dim)).to(self.device)
