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


def install_reference_aliases():
    for name in ("algorithms", "interfaces", "proposal_distributions", "target_distributions"):
        sys.modules[name] = sys.modules[__name__ + "." + name]


def build():
    from .build import build as _b
    return _b()
