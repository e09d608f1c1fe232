// Instantiates the fused kernel and the batched log-density kernel for the dense-covariance MultivariateNormal target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(mvn_dense, MVNDense)
