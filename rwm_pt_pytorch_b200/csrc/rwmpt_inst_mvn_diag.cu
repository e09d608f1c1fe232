// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the MVNDiag target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(mvn_diag, MVNDiag)
