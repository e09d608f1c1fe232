// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the ThreeMixture target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
// d = 50 (BASELINE config 4): 7 coordinates x 8 lanes -- two warps per 8-temperature ladder, ~130 registers; the
// 13 x 4 mapping (one warp per ladder) stays reachable with lanes_per_chain = 4.
#define TUNED_LIST(cls)                                              \
  RWMPT_TUNED_CASE_V(cls, 7, 8, 1, 5) RWMPT_TUNED_CASE_V(cls, 7, 8, 2, 5) \
  RWMPT_TUNED_CASE(cls, 7, 8, 1) RWMPT_TUNED_CASE(cls, 7, 8, 2)     \
  RWMPT_TUNED_CASE(cls, 13, 4, 1) RWMPT_TUNED_CASE(cls, 13, 4, 2)
RWMPT_DEFINE_TUNED(rwmpt::ThreeMixture, TUNED_LIST)
RWMPT_DEFINE_FAMILY(three_mixture, ThreeMixture)
