"""Distribution-level probe: mode occupancy of the cold chains of PT-RWM on RoughCarpet d=20 (weights .5/.3/.2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rwm_pt_pytorch_b200.algorithms import ParallelTemperingRWM_GPU_Optimized as PT, RandomWalkMH_GPU_Optimized as RWM
import rwm_pt_pytorch_b200.target_distributions as td
dev = torch.device("cuda", 0)
t = td.RoughCarpetDistributionTorch(20, device="cpu")
for mode in ("exchange", "reference"):
    for T in (20000, 100000, 400000):
        p = PT(20, 0.9, t, geom_temp_spacing=True, swap_every=10, burn_in=2000, device=dev, num_ladders=1024, seed=4, store="none", swap_mode=mode)
        p.generate_samples(T)
        x = p.current_states.view(1024, 8, 20)[:, 0].cpu().numpy()
        occ = [(x < -2.5).mean(), (np.abs(x) <= 2.5).mean(), (x > 2.5).mean()]
        print(mode, T, [round(float(o), 4) for o in occ], "swap", round(p.swap_acceptance_rate, 4), flush=True)
# Gaussian sanity (reference tests/test_pt_gpu_optimizations.py:91-93, tests/test_rwm_correctness.py:76-87)
g = td.MultivariateNormalTorch(8, device="cpu")
p = PT(8, 0.6, g, geom_temp_spacing=True, swap_every=10, burn_in=1000, device=dev, num_ladders=256, seed=2, store="cold", pre_allocate_steps=6000)
s = p.generate_samples(6000)
xs = s.reshape(-1, 8).double().cpu().numpy()
print("gauss8 mean err", np.abs(xs.mean(0)).max(), "cov err", np.abs(np.cov(xs.T) - np.eye(8)).max(), "swap", p.swap_acceptance_rate)
