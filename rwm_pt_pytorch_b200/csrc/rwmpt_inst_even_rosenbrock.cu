// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the EvenRosenbrock target.
#include "rwmpt_launch.cuh"
RWMPT_DEFINE_FAMILY(even_rosenbrock, EvenRosenbrock)
