// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the RoughCarpet target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
#include "rwmpt_spec.cuh"
#define TUNED_LIST(cls)                                               \
  RWMPT_TUNED_PLAIN_CASE(cls, rwmpt::RoughCarpetPlain, 5, 4, 0)       \
  RWMPT_TUNED_CASE(cls, 5, 4, 0)
RWMPT_DEFINE_TUNED(rwmpt::RoughCarpet, TUNED_LIST)
RWMPT_DEFINE_FAMILY(rough_carpet, RoughCarpet)
namespace rwmpt {
cudaError_t launch_mcmc_spec_rough_carpet_c3(const KernelArgs& a, int consumer_lanes, cudaStream_t st) {
  if (consumer_lanes == 1) return launch_mcmc_spec<RoughCarpetPlain, 5, 4, RWMPT_P_NORMAL, 1>(a, st);
  return launch_mcmc_spec<RoughCarpetPlain, 5, 4, RWMPT_P_NORMAL, 4>(a, st);
}
}  // namespace rwmpt
