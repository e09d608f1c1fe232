// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the IIDGamma target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(iid_gamma, IIDGamma)
