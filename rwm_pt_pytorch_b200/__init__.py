"""rwm_pt_pytorch_b200 -- B200-native (sm_100a) implementation of the RWM / PT-RWM sampling hot path of
aidanmrli/rwm-pt-pytorch, behind the reference's own Python classes.

    from rwm_pt_pytorch_b200.algorithms import RandomWalkMH_GPU_Optimized, ParallelTemperingRWM_GPU_Optimized
    from rwm_pt_pytorch_b200.interfaces import MCMCSimulation_GPU
    from rwm_pt_pytorch_b200.proposal_distributions import NormalProposal, LaplaceProposal, UniformRadiusProposal
    from rwm_pt_pytorch_b200.target_distributions import RoughCarpetDistributionTorch, ...

`install_reference_aliases()` additionally registers the four sub-packages under the reference's top-level
names (`algorithms`, `interfaces`, `proposal_distributions`, `target_distributions`) so that the reference's
driver scripts import this implementation unchanged.
"""
import sys

from . import _lib  # noqa: F401
from . import interfaces, proposal_distributions, target_distributions, algorithms  # noqa: F401

__version__ = "0.1.0"


_SUBPACKAGES = ("algorithms", "interfaces", "proposal_distributions", "target_distributions")


def install_reference_aliases(overlay=None):
    """Make the reference's import statements resolve to this implementation.

    overlay=False: register this package's four sub-packages under the reference's top-level names (`algorithms`,
    `interfaces`, `proposal_distributions`, `target_distributions`).  Enough for drivers that use only the GPU path.

    overlay=True: the REAL reference packages stay importable (its NumPy samplers `algorithms.rwm` / `algorithms.pt_rwm`,
    its NumPy targets, its plotting helpers are not part of this repository) and every class of the GPU hot path is
    replaced in place, on the reference's own modules, by the sm_100a-backed class of the same name -- what a maintainer
    switching the reference over would do.  The reference's own test scripts then run unmodified against this
    implementation (tests/test_reference_suite.py).

    overlay=None (default): overlay when the reference's packages are already imported or importable, else register.
    Returns the list of `module.attribute` names that were replaced (overlay) or the registered package names."""
    import importlib
    import importlib.util
    ours = {name: sys.modules[__name__ + "." + name] for name in _SUBPACKAGES}
    if overlay is None:
        def is_theirs(name):
            m = sys.modules.get(name)
            if m is not None:
                return m is not ours[name]
            try:
                return importlib.util.find_spec(name) is not None
            except (ImportError, ValueError):
                return False
        overlay = all(is_theirs(n) for n in _SUBPACKAGES)
    if not overlay:
        for name in _SUBPACKAGES:
            sys.modules[name] = ours[name]
        return list(_SUBPACKAGES)
    replaced = []
    for pkg in _SUBPACKAGES:
        theirs = importlib.import_module(pkg)
        if theirs is ours[pkg]:
            continue
        public = [n for n in getattr(ours[pkg], "__all__", []) if isinstance(getattr(ours[pkg], n), type)]
        mods = [theirs] + [m for k, m in list(sys.modules.items()) if k.startswith(pkg + ".") and m is not None]
        for cls_name in public:
            obj = getattr(ours[pkg], cls_name)
            for m in mods:
                if cls_name in vars(m) or m is theirs:
                    setattr(m, cls_name, obj)
                    replaced.append(f"{m.__name__}.{cls_name}")
    return replaced


def build():
    from .build import build as _b
    return _b()
