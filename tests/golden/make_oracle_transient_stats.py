"""NumPy-oracle run of the reference's RWM on EvenRosenbrock d=30 at the data/ sweep point x=0.085641 (variance x^2/d),
384 chains x 1e6 steps, burn-in 1000 -- about 10 minutes on one core.  Recorded in oracle_even_rosenbrock_d30.json:
the chains are still in their transient at 1e6 steps (the pooled acceptance drifts from 0.80 at 2e5 steps to 0.727 at
1e6), which is why the seed-averaged data/ value (0.7192 +- 0.0023 over ~25 single-chain files) and a 1e6-step run of
the reference algorithm itself differ by more than their standard errors.  tests/test_gpu_parity.py compares the kernel
with this run."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import rwmpt_oracle as O
from tests._util import product_target
t = product_target("even_rosenbrock_d30"); spec = t.spec()
d, x, B, T, burn = 30, 0.085641, 384, 1_000_000, 1000
rs = np.random.RandomState(5)
std = O.normal_std(x * x / d, 1.0)
X = (1e-8 * rs.randn(B, d)).astype(np.float32)
lp = O.log_density(spec, X).astype(np.float32)
acc_n = np.zeros(B); t0 = time.time()
for s in range(1, T + 1):
    inc = (rs.standard_normal((B, d)).astype(np.float32) * std)
    u = rs.random_sample(B).astype(np.float32)
    P = X + inc
    lpp = O.log_density(spec, P).astype(np.float32)
    a = O.accept_rule(lp, lpp, u, np.float32(1.0))[0]
    X = np.where(a[:, None], P, X); lp = np.where(a, lpp, lp)
    if s > burn: acc_n += a
    if s % 200000 == 0:
        r = acc_n / (s - burn); print(s, r.mean(), r.std(ddof=1) / np.sqrt(B), time.time() - t0, flush=True)
