"""Gaussian targets (reference: target_distributions/multivariate_normal_torch.py):
`MultivariateNormalTorch` (:5-134) -- identity / diagonal covariance runs the lean diagonal functor (`MVNDiag`), a general
dense covariance the gather-and-matvec functor (`MVNDense`, dim <= 128) -- and `ScaledMultivariateNormalTorch` (:137-295)."""
import numpy as np
import torch

from .. import _lib
from ..interfaces.target_torch import TorchTargetDistribution
from ._common import _MoveTensorsMixin, t2n


class MultivariateNormalTorch(_MoveTensorsMixin, TorchTargetDistribution):
    family_id = _lib.T_MVN_DIAG

    def __init__(self, dim, mean=None, cov=None, device=None):
        super().__init__(dim, device)
        self.name = "MultivariateNormalTorch"
        mean = torch.zeros(dim, dtype=torch.float32) if mean is None else torch.as_tensor(mean, dtype=torch.float32)
        cov = torch.eye(dim, dtype=torch.float32) if cov is None else torch.as_tensor(cov, dtype=torch.float32)
        self.is_diagonal = bool(torch.equal(torch.diag(torch.diagonal(cov)), cov))
        if not self.is_diagonal:
            if dim > 128:
                raise NotImplementedError("MultivariateNormalTorch with a dense covariance supports dim <= 128 "
                                          "(every lane of a chain gathers the whole state)")
            self.family_id = _lib.T_MVN_DENSE
        self.mean = mean.to(self.device)
        self.cov = cov.to(self.device)
        self.cov_inv = torch.linalg.inv(self.cov)
        self.cov_det = torch.linalg.det(self.cov)
        log_2pi = torch.log(torch.tensor(2.0 * torch.pi, device=self.device, dtype=torch.float32))
        self.log_norm_const = -0.5 * (dim * log_2pi + torch.log(self.cov_det))

    def _pack(self):
        if not self.is_diagonal:
            return torch.cat([self._header(float(self.log_norm_const)), self.mean.cpu(), self.cov_inv.cpu().reshape(-1)])
        return torch.cat([self._header(float(self.log_norm_const)), self.mean.cpu(), torch.diagonal(self.cov_inv).cpu()])

    def spec(self):
        if not self.is_diagonal:
            return dict(family="mvn_dense", mean=t2n(self.mean), cov_inv=t2n(self.cov_inv), log_norm_const=t2n(self.log_norm_const))
        return dict(family="mvn_diag", mean=t2n(self.mean), prec=t2n(torch.diagonal(self.cov_inv)),
                    log_norm_const=t2n(self.log_norm_const))

    def get_name(self):
        return self.name

    def draw_sample(self, beta=1.0):
        return np.random.multivariate_normal(self.mean.cpu().numpy(), self.cov.cpu().numpy() / beta)

    def draw_samples_torch(self, n_samples, beta=1.0):
        z = torch.randn(n_samples, self.dim, device=self.device, dtype=torch.float32)
        if not self.is_diagonal:                                           # :101-121: mean + z @ chol(cov / beta)^T
            return self.mean.unsqueeze(0) + torch.matmul(z, torch.linalg.cholesky(self.cov / beta).T)
        return self.mean.unsqueeze(0) + z * torch.sqrt(torch.diagonal(self.cov) / beta).unsqueeze(0)


class ScaledMultivariateNormalTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """pi(x) = prod_i c_i N(c_i x_i | 0, 1)."""
    family_id = _lib.T_SCALED_MVN

    def __init__(self, dim, scaling_factors=None, scaling_range=(0.02, 1.98), device=None, seed=None):
        super().__init__(dim, device)
        self.name = "ScaledMultivariateNormalTorch"
        if seed is not None:
            torch.manual_seed(seed)
        if scaling_factors is not None:
            self.scaling_factors = torch.as_tensor(scaling_factors).clone().detach().to(device=self.device, dtype=torch.float32)
        else:
            lo, hi = scaling_range
            self.scaling_factors = torch.rand(dim, device=self.device, dtype=torch.float32) * (hi - lo) + lo
        assert self.scaling_factors.shape == (dim,), f"Scaling factors must have shape ({dim},), got {self.scaling_factors.shape}"
        log_2pi = torch.log(torch.tensor(2.0 * torch.pi, device=self.device, dtype=torch.float32))
        self.log_norm_const = torch.sum(torch.log(self.scaling_factors)) - 0.5 * self.dim * log_2pi

    def _pack(self):
        return torch.cat([self._header(float(self.log_norm_const)), self.scaling_factors.cpu()])

    def spec(self):
        return dict(family="scaled_mvn", c=t2n(self.scaling_factors), log_norm_const=t2n(self.log_norm_const))

    def get_name(self):
        return self.name

    def draw_samples_torch(self, n_samples, beta=1.0):
        z = torch.randn(n_samples, self.dim, device=self.device, dtype=torch.float32)
        std = 1.0 / (self.scaling_factors * torch.sqrt(torch.tensor(beta, device=self.device, dtype=torch.float32)))
        return std.unsqueeze(0) * z

    def draw_sample(self, beta=1.0):
        return self.draw_samples_torch(1, beta)[0].cpu().numpy()

    def get_scaling_factors(self):
        return self.scaling_factors.clone()

    def get_variances(self):
        return 1.0 / (self.scaling_factors ** 2)
