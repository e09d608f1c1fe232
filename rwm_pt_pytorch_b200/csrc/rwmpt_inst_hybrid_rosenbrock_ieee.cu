// Parity-mode (IEEE arithmetic, injected randomness, decision outputs) instantiations of the fused kernel and the batched
// log-density kernel for the HybridRosenbrock target; the fast-math half is rwmpt_inst_hybrid_rosenbrock.cu.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY_IEEE(hybrid_rosenbrock, HybridRosenbrock)
