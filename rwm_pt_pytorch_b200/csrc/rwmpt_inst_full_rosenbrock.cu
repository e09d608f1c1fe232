// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the FullRosenbrock target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(full_rosenbrock, FullRosenbrock)
