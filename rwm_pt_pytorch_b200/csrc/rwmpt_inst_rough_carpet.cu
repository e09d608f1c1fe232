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
// warp-specialised kernel (rwmpt_spec.cuh): BASELINE config 3's shape, RoughCarpet without scaling block, 5 x 4, Normal
cudaError_t launch_spec_rough_carpet(const KernelArgs& a, int E, int W, int consumer_lanes, int producers, cudaStream_t st) {
  if (!a.target_plain || E != 5 || W != 4 || a.prop_family != RWMPT_P_NORMAL) return cudaErrorNotSupported;
  if (consumer_lanes == 1 && producers == 1) return launch_mcmc_spec<RoughCarpetPlain, 5, 4, RWMPT_P_NORMAL, 1, 1>(a, st);
  if (producers == 2) return launch_mcmc_spec<RoughCarpetPlain, 5, 4, RWMPT_P_NORMAL, 4, 2>(a, st);
  return launch_mcmc_spec<RoughCarpetPlain, 5, 4, RWMPT_P_NORMAL, 4, 1>(a, st);
}
}  // namespace rwmpt
