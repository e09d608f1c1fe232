// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the IIDBeta target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(iid_beta, IIDBeta)
