#!/usr/bin/env python3
"""Launches the stand-alone kernels once each (for ncu captures): ESJD reduction with and without the moved-row count,
batched log-density, proposal samplers, the swap-probability estimator of the ladder construction, the stand-alone sweep."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rwm_pt_pytorch_b200 import _lib  # noqa: E402
import rwm_pt_pytorch_b200.target_distributions as td  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
B, S, d = 4096, 1025, 50
x = torch.randn((B, S, d), device=dev)
x[:, 1::3] = x[:, 0:-1:3]                      # a third of the rows did not move
res = torch.empty(B, device=dev, dtype=torch.float64)
moved = torch.empty(B, device=dev, dtype=torch.int64)
for _ in range(3):
    _lib.check(lib.rwmpt_esjd_reduce(x.data_ptr(), B, S, 0, S, d, res.data_ptr(), None, _lib.stream_ptr(dev)))
    _lib.check(lib.rwmpt_esjd_reduce(x.data_ptr(), B, S, 0, S, d, res.data_ptr(), moved.data_ptr(), _lib.stream_ptr(dev)))
torch.cuda.synchronize()
print("moved fraction", float(moved.sum()) / (B * (S - 1)))
t = td.RoughCarpetDistributionTorch(20, device="cpu", mode_centers=[-15.0, 0.0, 15.0])
params = t.device_params(dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
for _ in range(3):
    _lib.check(lib.rwmpt_swap_prob_estimate(_lib.target_struct(t.family_id, 20, params), 1.0, 0.57, 4_000_000, 7, 0, acc.data_ptr(), _lib.stream_ptr(dev)))
torch.cuda.synchronize()
print("swap probability", float(acc) / 12_000_000)
