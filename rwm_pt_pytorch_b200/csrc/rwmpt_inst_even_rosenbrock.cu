// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the EvenRosenbrock target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
#define TUNED_LIST(cls)                                                                              \
  RWMPT_TUNED_CASE(cls, 5, 2, 0) RWMPT_TUNED_CASE(cls, 5, 4, 0) RWMPT_TUNED_CASE(cls, 8, 4, 0)
RWMPT_DEFINE_TUNED(rwmpt::EvenRosenbrock, TUNED_LIST)
RWMPT_DEFINE_FAMILY(even_rosenbrock, EvenRosenbrock)
