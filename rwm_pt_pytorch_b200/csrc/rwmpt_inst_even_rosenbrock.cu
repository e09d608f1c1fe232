// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the EvenRosenbrock target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
#include "rwmpt_spec.cuh"
#define TUNED_LIST(cls)                                                                              \
  RWMPT_TUNED_CASE(cls, 5, 2, 0) RWMPT_TUNED_CASE(cls, 5, 4, 0) RWMPT_TUNED_CASE(cls, 8, 4, 0)
RWMPT_DEFINE_TUNED(rwmpt::EvenRosenbrock, TUNED_LIST)
RWMPT_DEFINE_FAMILY(even_rosenbrock, EvenRosenbrock)
namespace rwmpt {
// warp-specialised kernel (rwmpt_spec.cuh): BASELINE config 2's exact shapes, d = 20 on 5 x 4 and d = 10 on 5 x 2, Normal
cudaError_t launch_spec_even_rosenbrock(const KernelArgs& a, int E, int W, int consumer_lanes, int producers, cudaStream_t st) {
  if (a.prop_family != RWMPT_P_NORMAL || E != 5 || a.dim != E * W) return cudaErrorNotSupported;
  (void)consumer_lanes;
  switch (W * 10 + producers) {
    case 41: return launch_mcmc_spec<EvenRosenbrock, 5, 4, RWMPT_P_NORMAL, 4, 1>(a, st);
    case 42: return launch_mcmc_spec<EvenRosenbrock, 5, 4, RWMPT_P_NORMAL, 4, 2>(a, st);
    case 43: return launch_mcmc_spec<EvenRosenbrock, 5, 4, RWMPT_P_NORMAL, 4, 3>(a, st);
    case 44: return launch_mcmc_spec<EvenRosenbrock, 5, 4, RWMPT_P_NORMAL, 4, 4>(a, st);
    case 21: return launch_mcmc_spec<EvenRosenbrock, 5, 2, RWMPT_P_NORMAL, 2, 1>(a, st);
    case 22: return launch_mcmc_spec<EvenRosenbrock, 5, 2, RWMPT_P_NORMAL, 2, 2>(a, st);
    case 23: return launch_mcmc_spec<EvenRosenbrock, 5, 2, RWMPT_P_NORMAL, 2, 3>(a, st);
    case 24: return launch_mcmc_spec<EvenRosenbrock, 5, 2, RWMPT_P_NORMAL, 2, 4>(a, st);
  }
  return cudaErrorNotSupported;
}
}  // namespace rwmpt
