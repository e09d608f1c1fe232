// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the IIDGamma target; the fast-math half is rwmpt_inst_iid_gamma.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(iid_gamma, IIDGamma)
