"""`NealFunnelTorch` (reference: target_distributions/funnel_torch.py:6-108) and `SuperFunnelTorch` (:111-348, hierarchical
logistic regression with per-group data: the data ride in the parameter block, every lane of a chain gathers the state)."""
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


class SuperFunnelTorch(_MoveTensorsMixin, TorchTargetDistribution):
    """Hierarchical logistic regression "super funnel" (reference: funnel_torch.py:111-348).  State
    theta = (alpha[J], beta[J*K], mu_alpha, mu_beta[K], tau_alpha, tau_beta); the per-group data (X_j, y_j) travel to the
    device inside the parameter block as N records (group, y, x[K]) (include/rwmpt.h RWMPT_T_SUPER_FUNNEL)."""
    family_id = _lib.T_SUPER_FUNNEL

    def __init__(self, J, K, X_data, Y_data, prior_hypermean_std=10.0, prior_tau_scale=2.5, device=None):
        self.J, self.K = int(J), int(K)
        dim = self.J + self.J * self.K + 1 + self.K + 1 + 1
        super().__init__(dim, device)
        if dim > 128:
            raise NotImplementedError("SuperFunnelTorch supports dim = J + J*K + K + 3 <= 128")
        if not (isinstance(X_data, list) and len(X_data) == J):
            raise ValueError(f"X_data must be a list of J={J} tensors.")
        if not (isinstance(Y_data, list) and len(Y_data) == J):
            raise ValueError(f"Y_data must be a list of J={J} tensors.")
        self.X_data, self.Y_data = [], []
        self.n_j_array = torch.empty(J, device=self.device, dtype=torch.long)
        for j in range(J):
            if not isinstance(X_data[j], torch.Tensor) or not isinstance(Y_data[j], torch.Tensor):
                raise ValueError(f"X_data[{j}] and Y_data[{j}] must be PyTorch tensors.")
            if X_data[j].ndim != 2 or X_data[j].shape[1] != K:
                raise ValueError(f"X_data[{j}] must have shape (n_j, K={K}). Got {X_data[j].shape}")
            if Y_data[j].ndim != 1 or Y_data[j].shape[0] != X_data[j].shape[0]:
                raise ValueError(f"Y_data[{j}] must have shape (n_j,). Got {Y_data[j].shape}, X_data had {X_data[j].shape[0]} samples.")
            self.X_data.append(X_data[j].to(self.device).to(torch.float32))
            self.Y_data.append(Y_data[j].to(self.device).to(torch.float32))
            self.n_j_array[j] = Y_data[j].shape[0]
        self.prior_hypermean_std = torch.tensor(prior_hypermean_std, device=self.device, dtype=torch.float32)
        self.prior_hypermean_var = self.prior_hypermean_std ** 2
        self.log_prior_hypermean_var = torch.log(self.prior_hypermean_var)
        self.prior_tau_scale = torch.tensor(prior_tau_scale, device=self.device, dtype=torch.float32)
        self.log_prior_tau_scale = torch.log(self.prior_tau_scale)
        self.log_2_pi = torch.tensor(2.0 * math.pi, device=self.device, dtype=torch.float32).log()
        self.log_2 = torch.tensor(2.0, device=self.device, dtype=torch.float32).log()
        self.log_pi = torch.tensor(math.pi, device=self.device, dtype=torch.float32).log()
        self.K_tensor = torch.tensor(float(self.K), device=self.device, dtype=torch.float32)

    def _records(self) -> torch.Tensor:
        recs = []
        for j in range(self.J):
            n = self.Y_data[j].shape[0]
            recs.append(torch.cat([torch.full((n, 1), float(j)), self.Y_data[j].cpu().reshape(n, 1), self.X_data[j].cpu()], dim=1))
        return torch.cat(recs, dim=0).to(torch.float32)

    def _pack(self):
        rec = self._records()
        head = self._header(float(self.J), float(self.K), float(self.prior_hypermean_var), float(self.log_prior_hypermean_var),
                            float(self.prior_tau_scale), float(self.log_prior_tau_scale), float(self.log_2_pi), float(self.log_2),
                            float(self.log_pi), float(rec.shape[0]))
        return torch.cat([head, rec.reshape(-1)])

    def spec(self):
        return dict(family="super_funnel", J=self.J, K=self.K, records=t2n(self._records()), hyper_var=t2n(self.prior_hypermean_var),
                    log_hyper_var=t2n(self.log_prior_hypermean_var), tau_scale=t2n(self.prior_tau_scale),
                    log_tau_scale=t2n(self.log_prior_tau_scale), log_2pi=t2n(self.log_2_pi), log_2=t2n(self.log_2), log_pi=t2n(self.log_pi))

    def get_name(self):
        return f"SuperFunnelTorch_J{self.J}_K{self.K}"

    def draw_sample(self, beta=1.0):
        raise NotImplementedError("SuperFunnelTorch.draw_sample is not implemented.")
