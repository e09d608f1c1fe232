// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the FullRosenbrock target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
#define TUNED_LIST(cls)                                             \
  RWMPT_TUNED_CASE_V(cls, 13, 8, 0, 1) RWMPT_TUNED_CASE_V(cls, 13, 8, 0, 2) RWMPT_TUNED_CASE(cls, 13, 8, 0)
RWMPT_DEFINE_TUNED(rwmpt::FullRosenbrock, TUNED_LIST)
RWMPT_DEFINE_FAMILY(full_rosenbrock, FullRosenbrock)
