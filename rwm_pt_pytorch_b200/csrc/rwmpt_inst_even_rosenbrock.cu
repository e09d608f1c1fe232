// Instantiates the fused RWM / PT-RWM kernel and the batched log-density kernel for the EvenRosenbrock target,
// plus the tuned (compile-time lanes-per-chain / proposal family) variants used by the BASELINE workloads.
#include "rwmpt_launch.cuh"
#include "rwmpt_spec.cuh"
#define TUNED_LIST(cls)                                                                              \
  RWMPT_TUNED_CASE(cls, 5, 2, 0) RWMPT_TUNED_CASE(cls, 5, 4, 0) RWMPT_TUNED_CASE(cls, 8, 4, 0)
RWMPT_DEFINE_TUNED(rwmpt::EvenRosenbrock, TUNED_LIST)
RWMPT_DEFINE_FAMILY(even_rosenbrock, EvenRosenbrock)
namespace rwmpt {
// warp-specialised kernel (rwmpt_spec.cuh): BASELINE config 2's shapes -- d = 20 on 5 x 4 and d = 10 on 5 x 2 (exact), d = 30 (any
// 24 < d <= 32) on 8 x 4 with masked padding coordinates -- Normal proposal
cudaError_t launch_spec_even_rosenbrock(const KernelArgs& a, int E, int W, int consumer_lanes, int producers, cudaStream_t st) {
  if (a.prop_family != RWMPT_P_NORMAL) return cudaErrorNotSupported;
  (void)consumer_lanes;
  if (E == 8 && W == 4 && a.dim > 24 && a.dim <= 32) {
    switch (producers) {
      case 1: return launch_mcmc_spec<EvenRosenbrock, 8, 4, RWMPT_P_NORMAL, 4, 1, false>(a, st);
      case 2: return launch_mcmc_spec<EvenRosenbrock, 8, 4, RWMPT_P_NORMAL, 4, 2, false>(a, st);
      case 3: return launch_mcmc_spec<EvenRosenbrock, 8, 4, RWMPT_P_NORMAL, 4, 3, false>(a, st);
      case 4: return launch_mcmc_spec<EvenRosenbrock, 8, 4, RWMPT_P_NORMAL, 4, 4, false>(a, st);
    }
    return cudaErrorNotSupported;
  }
  if (E == 4 && W == 8 && a.dim > 28 && a.dim <= 32) {
    switch (producers) {
      case 1: return launch_mcmc_spec<EvenRosenbrock, 4, 8, RWMPT_P_NORMAL, 8, 1, false>(a, st);
      case 2: return launch_mcmc_spec<EvenRosenbrock, 4, 8, RWMPT_P_NORMAL, 8, 2, false>(a, st);
      case 3: return launch_mcmc_spec<EvenRosenbrock, 4, 8, RWMPT_P_NORMAL, 8, 3, false>(a, st);
    }
    return cudaErrorNotSupported;
  }
  if (E != 5 || a.dim != E * W) return cudaErrorNotSupported;
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
