// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the ThreeMixture target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(three_mixture, ThreeMixture)
