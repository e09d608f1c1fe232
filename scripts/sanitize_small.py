"""Small invocations of every kernel path, meant to be run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
Covers: shuffle sweep (one warp per ladder), shared-memory sweep with CTA barriers (two warps per ladder, both swap
modes), staged trajectory stores + flush, balanced (ticketed) schedule, IEEE test mode with decisions, log-density,
proposal, stand-alone swap and both ESJD kernels, host-buffer entry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from rwm_pt_pytorch_b200 import _lib
from rwm_pt_pytorch_b200.algorithms import ParallelTemperingRWM_GPU_Optimized as PT, RandomWalkMH_GPU_Optimized as RWM
from rwm_pt_pytorch_b200.proposal_distributions import LaplaceProposal, UniformRadiusProposal, NormalProposal
import rwm_pt_pytorch_b200.target_distributions as td

dev = torch.device("cuda", 0)
lib = _lib.load()
rc = td.RoughCarpetDistributionTorch(20, device="cpu")
for mode in ("reference", "exchange"):
    for store in ("none", "cold", "all"):
        p = PT(20, 0.9, rc, geom_temp_spacing=True, swap_every=4, burn_in=10, device=dev, num_ladders=5, seed=1, store=store,
               swap_mode=mode, pre_allocate_steps=90 if store != "none" else None)
        p.generate_samples(80)
d = 50
tm = td.ThreeMixtureDistributionTorch(d, device="cpu", mode_centers=[[-15.0] + [0.0] * (d - 1), [0.0] * d, [15.0] + [0.0] * (d - 1)])
for prop in (LaplaceProposal(d, torch.full((d,), 0.1), 1.0, torch.device("cpu"), torch.float32),
             UniformRadiusProposal(d, 1.0, 1.0, torch.device("cpu"), torch.float32)):
    for mode in ("reference", "exchange"):
        p = PT(d, None, tm, geom_temp_spacing=True, swap_every=5, burn_in=0, device=dev, num_ladders=3, seed=2, store="all",
               swap_mode=mode, proposal_distribution=prop, pre_allocate_steps=70, initial_states=np.zeros((3, 8, d), np.float32))
        p.generate_samples(70)
        p.expected_squared_jump_distance_gpu()
# balanced schedule forced, RWM with a partial CTA, thinned stores
r = RWM(20, 0.4, rc, burn_in=7, device=dev, num_chains=37, seed=3, store="all", thin=3, pre_allocate_steps=200)
r._ensure_batch(1); r._batch.schedule = 2; r._batch.run(151); r._refresh_stats()
q = PT(20, 0.9, rc, geom_temp_spacing=True, swap_every=10, burn_in=20, device=dev, num_ladders=9, seed=5, store="none")
q._require_batch().schedule = 2
q.generate_samples(300)
# IEEE test mode with injected randomness and decision outputs
rs = np.random.RandomState(0)
T, L, K = 30, 3, 8
a = PT(20, 0.9, rc, geom_temp_spacing=True, swap_every=5, device=dev, num_ladders=L, store="none", math_mode="ieee",
       initial_states=np.zeros((L, K, 20), np.float32))
a.run_injected(rs.randn(T, L * K, 20).astype(np.float32), rs.rand(T, L * K).astype(np.float32), rs.rand(T // 5, L, K - 1).astype(np.float32))
# stand-alone kernels
x = torch.randn(1000, 20, device=dev)
rc.to(dev).log_density(x)
NormalProposal(20, 0.5, 1.0, dev, torch.float32).sample(257)
LaplaceProposal(20, torch.full((20,), 0.5), 1.0, dev, torch.float32).sample(257)
UniformRadiusProposal(20, 1.0, 1.0, dev, torch.float32).sample(257)
for (B, S, dd) in [(5, 400, 20), (3, 77, 7), (4, 300, 50)]:
    xs = torch.randn(B, S, dd, device=dev)
    out = torch.empty(B, device=dev, dtype=torch.float64); mv = torch.empty(B, device=dev, dtype=torch.int64)
    _lib.check(lib.rwmpt_esjd_reduce(xs.data_ptr(), B, S, 1, S - 1, dd, out.data_ptr(), None, _lib.stream_ptr(dev)))
    _lib.check(lib.rwmpt_esjd_reduce(xs.data_ptr(), B, S, 1, S - 1, dd, out.data_ptr(), mv.data_ptr(), _lib.stream_ptr(dev)))
torch.cuda.synchronize()
print("sanitize_small: all paths ran")
