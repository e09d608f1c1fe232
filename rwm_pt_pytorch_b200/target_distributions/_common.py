"""Helpers shared by the target parameter carriers."""
import torch


class _MoveTensorsMixin:
    """`.to(device)` moves every tensor attribute, like the reference's per-class `to` methods."""

    def to(self, device):
        device = torch.device(device)
        self.device = device
        for k, v in list(vars(self).items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


def t2n(t):
    return t.detach().cpu().numpy()
