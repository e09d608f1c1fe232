"""NumPy-oracle runs of the reference's RWM on EvenRosenbrock at the three data/ sweep points the GPU parity tests use
(variance x^2/d): 384 chains x 1e6 steps, burn-in 1000 -- about 10 minutes per case on one core.

    python tests/golden/make_oracle_transient_stats.py [d ...]        # default: 10 20 30

Recorded in oracle_even_rosenbrock_d{10,20,30}.json (pooled acceptance and ESJD with their standard errors at every
2e5 steps).  Why they exist: the chains are still in their transient at 1e6 steps (d=30: the pooled acceptance drifts
from 0.80 at 2e5 steps to 0.727 at 1e6), and the reference's recorded "seeds" share one random stream (SURVEY.md
section 0, item 2: the CUDA generator is never seeded), so the seed-averaged data/ values carry a much larger
uncertainty than their quoted standard errors.  tests/test_gpu_parity.py compares the kernel with these runs of the
reference ALGORITHM (same start, same length) within 3 standard errors, no slack."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from oracle import rwmpt_oracle as O  # noqa: E402
from tests._util import product_target  # noqa: E402

POINTS = {10: 0.161282, 20: 0.297436, 30: 0.085641}   # the x of tests/test_gpu_parity.py's data/ points
HERE = os.path.dirname(os.path.abspath(__file__))


def run(d, B=384, T=1_000_000, burn=1000, seed=5):
    x = POINTS[d]
    key = f"even_rosenbrock_d{d}"
    spec = product_target(key).spec()
    rs = np.random.RandomState(seed)
    std = O.normal_std(x * x / d, 1.0)
    X = (1e-8 * rs.randn(B, d)).astype(np.float32)
    lp = O.log_density(spec, X).astype(np.float32)
    acc_n = np.zeros(B)
    sq = np.zeros(B)
    rec = {"config": {"target": key, "x": x, "chains": B, "steps": T, "burn_in": burn, "seed": seed},
           "acceptance_at_steps": {}, "esjd_at_steps": {}}
    t0 = time.time()
    for s in range(1, T + 1):
        inc = (rs.standard_normal((B, d)).astype(np.float32) * std)
        u = rs.random_sample(B).astype(np.float32)
        P = X + inc
        lpp = O.log_density(spec, P).astype(np.float32)
        a = O.accept_rule(lp, lpp, u, np.float32(1.0))[0]
        if s > burn:
            acc_n += a
            dx = (np.where(a[:, None], P, X) - X).astype(np.float64)
            sq += (dx * dx).sum(1)
        X = np.where(a[:, None], P, X)
        lp = np.where(a, lpp, lp)
        if s % 200000 == 0:
            r, e = acc_n / (s - burn), sq / (s - burn)
            rec["acceptance_at_steps"][str(s)] = [float(r.mean()), float(r.std(ddof=1) / np.sqrt(B))]
            rec["esjd_at_steps"][str(s)] = [float(e.mean()), float(e.std(ddof=1) / np.sqrt(B))]
            rec["per_chain_std_at_%d" % s] = {"acceptance": float(r.std(ddof=1)), "esjd": float(e.std(ddof=1))}
            print(d, s, r.mean(), e.mean(), time.time() - t0, flush=True)
    with open(os.path.join(HERE, f"oracle_{key}.json"), "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    for d in ([int(a) for a in sys.argv[1:]] or [10, 20, 30]):
        run(d)
