// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the HybridRosenbrock target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(hybrid_rosenbrock, HybridRosenbrock)
