from .multimodal_torch import ThreeMixtureDistributionTorch, RoughCarpetDistributionTorch
from .rosenbrock_torch import FullRosenbrockTorch, EvenRosenbrockTorch, HybridRosenbrockTorch
from .funnel_torch import NealFunnelTorch, SuperFunnelTorch
from .hypercube_torch import HypercubeTorch
from .iid_product_torch import IIDGammaTorch, IIDBetaTorch
from .multivariate_normal_torch import MultivariateNormalTorch, ScaledMultivariateNormalTorch

__all__ = ["ThreeMixtureDistributionTorch", "RoughCarpetDistributionTorch", "FullRosenbrockTorch",
           "EvenRosenbrockTorch", "HybridRosenbrockTorch", "NealFunnelTorch", "SuperFunnelTorch", "HypercubeTorch",
           "IIDGammaTorch", "IIDBetaTorch", "MultivariateNormalTorch", "ScaledMultivariateNormalTorch"]
