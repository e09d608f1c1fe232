// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the FullRosenbrock target; the fast-math half is rwmpt_inst_full_rosenbrock.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(full_rosenbrock, FullRosenbrock)
