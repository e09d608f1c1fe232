// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the SuperFunnel target; the fast-math half is rwmpt_inst_super_funnel.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(super_funnel, SuperFunnel)
