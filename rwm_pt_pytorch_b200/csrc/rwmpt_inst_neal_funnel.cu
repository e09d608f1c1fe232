// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the NealFunnel target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(neal_funnel, NealFunnel)
