// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the ScaledMVN target; the fast-math half is rwmpt_inst_scaled_mvn.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(scaled_mvn, ScaledMVN)
