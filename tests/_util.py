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
