// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the RoughCarpet target; the fast-math half is rwmpt_inst_rough_carpet.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(rough_carpet, RoughCarpet)
