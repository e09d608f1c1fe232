"""A/B of the native-RNG statistics of variant libraries on the reference's EvenRosenbrock sweep points.
   RWMPT_LIB=<lib> python scripts/gpu_stat_ab.py [chains] [steps]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import product_target
from rwm_pt_pytorch_b200.algorithms import RandomWalkMH_GPU_Optimized as RWM
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
for key, d, x, lanes in [("even_rosenbrock_d30", 30, 0.085641, 0), ("even_rosenbrock_d30", 30, 0.085641, 4), ("even_rosenbrock_d20", 20, 0.297436, 0), ("even_rosenbrock_d10", 10, 0.161282, 0)]:
    for seed in (12345, 777):
        np.random.seed(7)
        algo = RWM(d, x * x / d, product_target(key), burn_in=1000, device="cuda", num_chains=B, seed=seed, lanes_per_chain=lanes)
        algo.generate_samples(T)
        acc = algo.acceptance_rates.cpu().numpy(); esjd = algo.esjd_per_chain().cpu().numpy()
        print(f"{key} lanes={lanes} seed={seed}: acc {acc.mean():.5f} +- {acc.std(ddof=1)/np.sqrt(B):.5f}   esjd {esjd.mean():.6f} +- {esjd.std(ddof=1)/np.sqrt(B):.6f}", flush=True)
