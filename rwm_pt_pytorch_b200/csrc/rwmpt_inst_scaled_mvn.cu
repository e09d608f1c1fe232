// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the ScaledMVN target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(scaled_mvn, ScaledMVN)
