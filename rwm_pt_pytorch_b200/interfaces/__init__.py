from .metropolis import MHAlgorithm
from .target import TargetDistribution
from .target_torch import TorchTargetDistribution
from .simulation_gpu import MCMCSimulation_GPU

__all__ = ["MHAlgorithm", "TargetDistribution", "TorchTargetDistribution", "MCMCSimulation_GPU"]
