// Instantiates the fused kernel and the batched log-density kernel for the SuperFunnel (hierarchical logistic) target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(super_funnel, SuperFunnel)
