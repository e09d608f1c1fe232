// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the RoughCarpet target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
#define TUNED_LIST(cls)                                               \
  RWMPT_TUNED_PLAIN_CASE_V(cls, rwmpt::RoughCarpetPlain, 5, 4, 0, 1)  \
  RWMPT_TUNED_PLAIN_CASE_V(cls, rwmpt::RoughCarpetPlain, 5, 4, 0, 2)  \
  RWMPT_TUNED_PLAIN_CASE(cls, rwmpt::RoughCarpetPlain, 5, 4, 0)       \
  RWMPT_TUNED_CASE(cls, 5, 4, 0)
RWMPT_DEFINE_TUNED(rwmpt::RoughCarpet, TUNED_LIST)
RWMPT_DEFINE_FAMILY(rough_carpet, RoughCarpet)
