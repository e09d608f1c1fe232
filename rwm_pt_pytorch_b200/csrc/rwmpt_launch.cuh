// rwmpt_launch.cuh -- host-side launch helpers; each rwmpt_inst_<family>.cu instantiates one target family
// for every supported elements-per-lane count, so the families compile in parallel.
#pragma once

#include "rwmpt_kernel.cuh"

namespace rwmpt {

// elements-per-lane (E) variants compiled for each math mode; the picker in rwmpt_api.cu uses the same lists
#define RWMPT_FAST_E_LIST(X) X(1) X(2) X(3) X(4) X(5) X(8) X(13)
#define RWMPT_IEEE_E_LIST(X) X(1) X(2) X(3) X(4) X(5) X(8) X(13)

template <template <int, bool> class Target, int E, bool IEEE, int WT, int PF, bool EXACT, bool TEST>
cudaError_t launch_mcmc_one(const KernelArgs& a, const LaunchGeom& g, cudaStream_t st) {
  auto kern = mcmc_kernel<Target, E, IEEE, WT, PF, EXACT, TEST>;
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)g.grid, g.threads, g.smem, st>>>(a);
  return cudaGetLastError();
}

template <template <int, bool> class Target, int E, bool IEEE>
cudaError_t launch_logp_one(const float* P, int d, int W, const float* x, long long n, float* out, cudaStream_t st) {
  const int threads = 128;
  const int per_cta = threads / W;
  long long blocks = (n + per_cta - 1) / per_cta;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  logp_kernel<Target, E, IEEE><<<(unsigned)blocks, threads, 0, st>>>(P, d, W, x, n, out);
  return cudaGetLastError();
}

// Tuned variants (lanes-per-chain and proposal family fixed at compile time) exist for the BASELINE workloads; each
// family's translation unit lists its own in Tuned<Target>::launch and returns cudaErrorNotSupported otherwise.
template <template <int, bool> class Target>
struct Tuned {
  static cudaError_t launch(const KernelArgs&, const LaunchGeom&, cudaStream_t) { return cudaErrorNotSupported; }
};

// generic variants: runtime lanes-per-chain and proposal family.  IEEE kernels carry the test mode (injection,
// decision outputs); fast kernels do not.
template <template <int, bool> class Target>
cudaError_t launch_mcmc_family(const KernelArgs& a, const LaunchGeom& g, bool ieee, cudaStream_t st) {
  if (ieee) {
    switch (g.E) {
#define X(e) case e: return launch_mcmc_one<Target, e, true, 0, -1, false, true>(a, g, st);
      RWMPT_IEEE_E_LIST(X)
#undef X
    }
  } else {
    cudaError_t e = Tuned<Target>::launch(a, g, st);
    if (e != cudaErrorNotSupported) return e;
    switch (g.E) {
#define X(e) case e: return launch_mcmc_one<Target, e, false, 0, -1, false, false>(a, g, st);
      RWMPT_FAST_E_LIST(X)
#undef X
    }
  }
  return cudaErrorInvalidValue;
}

template <template <int, bool> class Target>
cudaError_t launch_logp_family(const float* P, int d, int E, int W, const float* x, long long n, float* out, bool ieee,
                               cudaStream_t st) {
  if (ieee) {
    switch (E) {
#define X(e) case e: return launch_logp_one<Target, e, true>(P, d, W, x, n, out, st);
      RWMPT_IEEE_E_LIST(X)
#undef X
    }
  } else {
    switch (E) {
#define X(e) case e: return launch_logp_one<Target, e, false>(P, d, W, x, n, out, st);
      RWMPT_FAST_E_LIST(X)
#undef X
    }
  }
  return cudaErrorInvalidValue;
}

// one pair of entry points per family, defined in rwmpt_inst_<family>.cu
#define RWMPT_FAMILY_LIST(X)                 \
  X(rough_carpet, RoughCarpet)               \
  X(three_mixture, ThreeMixture)             \
  X(full_rosenbrock, FullRosenbrock)         \
  X(even_rosenbrock, EvenRosenbrock)         \
  X(hybrid_rosenbrock, HybridRosenbrock)     \
  X(neal_funnel, NealFunnel)                 \
  X(hypercube, Hypercube)                    \
  X(iid_gamma, IIDGamma)                     \
  X(iid_beta, IIDBeta)                       \
  X(scaled_mvn, ScaledMVN)                   \
  X(mvn_diag, MVNDiag)

#define X(name, cls)                                                                                         \
  cudaError_t launch_mcmc_##name(const KernelArgs& a, const LaunchGeom& g, bool ieee, cudaStream_t st);       \
  cudaError_t launch_logp_##name(const float* P, int d, int E, int W, const float* x, long long n, float* out, \
                                 bool ieee, cudaStream_t st);
RWMPT_FAMILY_LIST(X)
#undef X

// RWMPT_DEFINE_TUNED(cls, LIST) with LIST(X) = X(E, W, PF) ... specialises Tuned<cls>
#define RWMPT_TUNED_CASE(cls, e, w, pf)                                                        \
  if (g.E == e && g.W == w && a.prop_family == pf) {                              \
    if (a.dim == e * w) return launch_mcmc_one<cls, e, false, w, pf, true, false>(a, g, st);  \
    return launch_mcmc_one<cls, e, false, w, pf, false, false>(a, g, st);                     \
  }
// tuned case that swaps in a leaner functor (e.g. RoughCarpetPlain) when the target carries no scaling block
#define RWMPT_TUNED_PLAIN_CASE(cls, plain, e, w, pf)                                               \
  if (a.target_plain && g.E == e && g.W == w && a.prop_family == pf) {                             \
    if (a.dim == e * w) return launch_mcmc_one<plain, e, false, w, pf, true, false>(a, g, st);    \
    return launch_mcmc_one<plain, e, false, w, pf, false, false>(a, g, st);                       \
  }
#define RWMPT_DEFINE_TUNED(cls, LIST)                                                              \
  namespace rwmpt {                                                                                \
  template <>                                                                                      \
  struct Tuned<cls> {                                                                              \
    static cudaError_t launch(const KernelArgs& a, const LaunchGeom& g, cudaStream_t st) {         \
      LIST(cls)                                                                                    \
      return cudaErrorNotSupported;                                                                \
    }                                                                                              \
  };                                                                                               \
  }

#define RWMPT_DEFINE_FAMILY(name, cls)                                                                        \
  namespace rwmpt {                                                                                           \
  cudaError_t launch_mcmc_##name(const KernelArgs& a, const LaunchGeom& g, bool ieee, cudaStream_t st) {      \
    return launch_mcmc_family<cls>(a, g, ieee, st);                                                           \
  }                                                                                                           \
  cudaError_t launch_logp_##name(const float* P, int d, int E, int W, const float* x, long long n, float* out, \
                                 bool ieee, cudaStream_t st) {                                                \
    return launch_logp_family<cls>(P, d, E, W, x, n, out, ieee, st);                                          \
  }                                                                                                           \
  }

}  // namespace rwmpt
