from .rwm_gpu_optimized import RandomWalkMH_GPU_Optimized, RandomWalkMetropolis
from .pt_rwm_gpu_optimized import ParallelTemperingRWM_GPU_Optimized

__all__ = ["RandomWalkMH_GPU_Optimized", "RandomWalkMetropolis", "ParallelTemperingRWM_GPU_Optimized"]
