"""Shared helpers for the test-suite: golden-fixture loading and spec handling."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """Return (spec, arrays) from tests/golden/<name>.npz (written by tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    spec, arrays = {}, {}
    for k in z.files:
        v = z[k]
        if k.startswith("spec_"):
            kk = k[5:]
            if kk == "family":
                spec[kk] = str(v)
            elif v.ndim == 0:
                spec[kk] = v.item() if v.dtype.kind in "iu" else np.float32(v)
            else:
                spec[kk] = v
        else:
            arrays[k] = v
    return spec, arrays


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


# ---- Philox4x32-10 in pure Python (test-side cross-check of the known-answer vectors) ----------------
def philox4x32_10_py(ctr, key):
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return [c0, c1, c2, c3]


# Random123 kat_vectors, philox4x32 10 rounds
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


# ---- product target objects --------------------------------------------------------------------------
def make_product_targets():
    """Same constructor calls, in the same order and under the same torch seed, as make_targets() of
    tests/golden/make_golden.py -- so the random scaling factors coincide with the reference run."""
    import torch
    import rwm_pt_pytorch_b200.target_distributions as td
    CPU = torch.device("cpu")
    torch.manual_seed(1234)
    d = {}
    d["rough_carpet_d20"] = td.RoughCarpetDistributionTorch(20, device=CPU)
    d["rough_carpet_pm4_d20"] = td.RoughCarpetDistributionTorch(20, device=CPU, mode_centers=[-4.0, 0.0, 4.0])
    d["rough_carpet_scaled_d6"] = td.RoughCarpetDistributionTorch(6, scaling=True, device=CPU)
    d["three_mixture_d10"] = td.ThreeMixtureDistributionTorch(10, device=CPU)
    d["three_mixture_pm15_d50"] = td.ThreeMixtureDistributionTorch(
        50, device=CPU, mode_centers=[[-15.0] + [0.0] * 49, [0.0] * 50, [15.0] + [0.0] * 49])
    d["three_mixture_scaled_d7"] = td.ThreeMixtureDistributionTorch(7, scaling=True, device=CPU)
    d["full_rosenbrock_d20"] = td.FullRosenbrockTorch(20, device=CPU)
    d["full_rosenbrock_d3"] = td.FullRosenbrockTorch(3, device=CPU)
    d["even_rosenbrock_d10"] = td.EvenRosenbrockTorch(10, device=CPU)
    d["even_rosenbrock_d20"] = td.EvenRosenbrockTorch(20, device=CPU)
    d["even_rosenbrock_d30"] = td.EvenRosenbrockTorch(30, device=CPU)
    d["hybrid_rosenbrock_n3x5"] = td.HybridRosenbrockTorch(3, 5, device=CPU)
    d["hybrid_rosenbrock_n4x2"] = td.HybridRosenbrockTorch(4, 2, device=CPU)
    d["neal_funnel_d10"] = td.NealFunnelTorch(10, device=CPU)
    d["neal_funnel_d1"] = td.NealFunnelTorch(1, device=CPU)
    d["hypercube_pm1_d5"] = td.HypercubeTorch(5, left_boundary=-1, right_boundary=1, device=CPU)
    d["hypercube_01_d4"] = td.HypercubeTorch(4, device=CPU)
    d["iid_gamma_d8"] = td.IIDGammaTorch(8, shape=2, scale=3, device=CPU)
    d["iid_beta_d8"] = td.IIDBetaTorch(8, alpha=2, beta=3, device=CPU)
    d["scaled_mvn_d12"] = td.ScaledMultivariateNormalTorch(12, device=CPU)
    d["mvn_identity_d50"] = td.MultivariateNormalTorch(50, device=CPU)
    d["mvn_diag_d6"] = td.MultivariateNormalTorch(
        6, mean=[0.5, -1.0, 0.0, 2.0, 0.25, -0.75], cov=np.diag([0.5, 2.0, 1.0, 4.0, 0.25, 1.5]).tolist(), device=CPU)
    d["full_rosenbrock_d100"] = td.FullRosenbrockTorch(100, device=CPU)
    d["neal_funnel_d100"] = td.NealFunnelTorch(100, device=CPU)
    dense_mean = [0.3 * ((-1) ** i) * (i % 4) for i in range(12)]
    dense_cov = [[(0.7 ** abs(i - j)) * (1.0 + 0.1 * min(i, j)) for j in range(12)] for i in range(12)]
    d["mvn_dense_d12"] = td.MultivariateNormalTorch(12, mean=dense_mean, cov=dense_cov, device=CPU)
    g = torch.Generator().manual_seed(42)
    X, Y = [], []
    for _ in range(5):
        Xj = torch.randn(20, 3, generator=g)
        X.append(Xj)
        Y.append(torch.bernoulli(torch.sigmoid(0.5 * torch.sum(Xj, dim=1)), generator=g))
    d["super_funnel_j5k3"] = td.SuperFunnelTorch(5, 3, X, Y, prior_hypermean_std=10.0, prior_tau_scale=2.5, device=CPU)
    return d


_PRODUCT_TARGETS = None


def product_target(key):
    global _PRODUCT_TARGETS
    if _PRODUCT_TARGETS is None:
        _PRODUCT_TARGETS = make_product_targets()
    return _PRODUCT_TARGETS[key]


# golden file stem -> key in make_product_targets()
def target_key_of(golden_name):
    for prefix in ("logp_", "rwm_", "pt_"):
        if golden_name.startswith(prefix):
            rest = golden_name[len(prefix):]
            keys = sorted(make_keys(), key=len, reverse=True)
            for k in keys:
                if rest.startswith(k):
                    return k
    raise KeyError(golden_name)


def make_keys():
    return ["rough_carpet_d20", "rough_carpet_pm4_d20", "rough_carpet_scaled_d6", "three_mixture_d10",
            "three_mixture_pm15_d50", "three_mixture_scaled_d7", "full_rosenbrock_d20", "full_rosenbrock_d3",
            "even_rosenbrock_d10", "even_rosenbrock_d20", "even_rosenbrock_d30", "hybrid_rosenbrock_n3x5",
            "hybrid_rosenbrock_n4x2", "neal_funnel_d10", "neal_funnel_d1", "hypercube_pm1_d5", "hypercube_01_d4",
            "iid_gamma_d8", "iid_beta_d8", "scaled_mvn_d12", "mvn_identity_d50", "mvn_diag_d6", "full_rosenbrock_d100",
            "neal_funnel_d100", "mvn_dense_d12", "super_funnel_j5k3"]


def specs_equal(a, b):
    if set(a) != set(b):
        return False
    for k in a:
        if k == "family":
            if a[k] != b[k]:
                return False
        elif not np.array_equal(np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)):
            return False
    return True


def near_tie_report(dec_gpu, dec_ref, lar_ref, u, rel_tol):
    """Compare accept decisions chain by chain.  A flipped decision cascades, so each chain is compared up to
    its first mismatch only; a mismatch is a *near tie* when |u - exp(lar)| <= rel_tol * exp(lar) at that step
    (fp32 reductions in a different order move lar by an ulp).  Returns (n_near_ties, n_hard_mismatches,
    valid_until[c])."""
    dec_gpu = np.asarray(dec_gpu).astype(bool)
    dec_ref = np.asarray(dec_ref).astype(bool)
    T, B = dec_ref.shape
    valid_until = np.full(B, T, dtype=np.int64)
    near = hard = 0
    for c in range(B):
        mism = np.nonzero(dec_gpu[:, c] != dec_ref[:, c])[0]
        if mism.size == 0:
            continue
        t = int(mism[0])
        valid_until[c] = t
        with np.errstate(over="ignore"):
            p = float(np.exp(np.float64(lar_ref[t, c])))
        if abs(float(u[t, c]) - p) <= rel_tol * max(p, 1e-30):
            near += 1
        else:
            hard += 1
    return near, hard, valid_until
