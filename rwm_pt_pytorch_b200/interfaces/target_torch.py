"""`TorchTargetDistribution` base (reference: interfaces/target_torch.py:5-67).

Same public surface as the reference (`dim`, `device`, `density`, `log_density`, `get_name`, `draw_sample`,
`to`).  A target here is a *parameter carrier*: `family_id` + `pack()` give the flat float32 parameter block
of include/rwmpt.h, and `log_density` runs the hand-written CUDA functor (`rwmpt_log_density`).  There is
no CPU evaluation path.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from .. import _lib


class TorchTargetDistribution(ABC):
    family_id: int = -1

    def __init__(self, dimension, device=None):
        self.dim = dimension
        if device is None:
            self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        else:
            self.device = torch.device(device)
        self._packed = None
        self._packed_dev = {}
        self.math_mode = "fast"   # "ieee" = parity mode, reference operation order

    # ---- parameter block for the C ABI --------------------------------------------------------
    @abstractmethod
    def _pack(self) -> torch.Tensor:
        """Flat float32 CPU tensor: 16 scalars then per-dimension vectors (layout in include/rwmpt.h)."""

    def pack(self) -> torch.Tensor:
        if self._packed is None:
            p = self._pack().detach().to(device="cpu", dtype=torch.float32).contiguous()
            assert p.numel() >= _lib.PARAM_HEADER
            self._packed = p
        return self._packed

    def device_params(self, device) -> torch.Tensor:
        device = torch.device(device)
        key = (device.type, device.index)
        if key not in self._packed_dev:
            self._packed_dev[key] = self.pack().to(device)
        return self._packed_dev[key]

    def _invalidate(self):
        self._packed = None
        self._packed_dev = {}

    @staticmethod
    def _header(*scalars) -> torch.Tensor:
        h = torch.zeros(_lib.PARAM_HEADER, dtype=torch.float32)
        for i, s in enumerate(scalars):
            h[i] = float(s)
        return h

    # ---- reference API --------------------------------------------------------------------------
    def log_density(self, x):
        """log pi(x) for x of shape (dim,) -> () or (batch, dim) -> (batch,), evaluated on the GPU."""
        dev = _lib.require_cuda(self.device if torch.device(self.device).type == "cuda" else "cuda")
        lib = _lib.load()
        x = torch.as_tensor(x)
        single = x.ndim == 1
        if x.ndim not in (1, 2) or x.shape[-1] != self.dim:
            raise ValueError(f"Expected tensor of shape ({self.dim},) or (batch_size, {self.dim}), got {tuple(x.shape)}")
        xb = x.detach().to(device=dev, dtype=torch.float32).reshape(-1, self.dim).contiguous()
        out = torch.empty(xb.shape[0], device=dev, dtype=torch.float32)
        params = self.device_params(dev)
        tgt = _lib.target_struct(self.family_id, self.dim, params)
        with torch.cuda.device(dev):
            _lib.check(lib.rwmpt_log_density(tgt, xb.data_ptr(), xb.shape[0], out.data_ptr(),
                                             _lib.MATH_MODES[self.math_mode], _lib.stream_ptr(dev)))
        return out[0] if single else out

    def density(self, x):
        return torch.exp(self.log_density(x))

    @abstractmethod
    def get_name(self):
        raise NotImplementedError("Subclasses must implement the get_name method.")

    def draw_sample(self, beta=1.0):
        raise NotImplementedError("Subclasses should implement draw_sample for compatibility.")

    def to(self, device):
        self.device = torch.device(device)
        return self
