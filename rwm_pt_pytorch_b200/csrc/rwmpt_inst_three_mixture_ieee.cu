// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the ThreeMixture target; the fast-math half is rwmpt_inst_three_mixture.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(three_mixture, ThreeMixture)
