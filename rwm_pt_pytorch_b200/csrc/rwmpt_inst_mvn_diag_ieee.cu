// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the MVNDiag target; the fast-math half is rwmpt_inst_mvn_diag.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(mvn_diag, MVNDiag)
