// rwmpt_launch.cuh -- host-side launch helpers; each rwmpt_inst_<family>.cu instantiates one target family
// for every supported elements-per-lane count, so the families compile in parallel.
#pragma once

#include "rwmpt_kernel.cuh"

namespace rwmpt {

// elements-per-lane (E) variants compiled for each math mode; the picker in rwmpt_api.cu uses the same lists
#ifdef RWMPT_ONLY_TUNED  // analysis builds (SASS inspection of one kernel): a single generic variant
#define RWMPT_FAST_E_LIST(X) X(5)
#define RWMPT_IEEE_E_LIST(X) X(5)
#else
// (25: few lanes per chain for LONG ladders -- a ladder's n_temps * lanes threads must fit one CTA of kMaxCtaThreads, so
// d = 100 runs ladders of up to 64 temperatures on 25 coordinates x 4 lanes -- and dimensions up to 800)
#define RWMPT_FAST_E_LIST(X) X(1) X(2) X(3) X(4) X(5) X(8) X(13) X(25)
#define RWMPT_IEEE_E_LIST(X) X(1) X(2) X(3) X(4) X(5) X(8) X(13) X(25)
#endif

template <template <int, bool> class Target, int E, bool IEEE, int WT, int PF, bool EXACT, bool TEST, bool STORE, int VARIANT = 0>
cudaError_t launch_mcmc_store(const KernelArgs& a_in, const LaunchGeom& g, cudaStream_t st);

// Fast (non-test) kernels exist twice: with and without the retained-sample path, so that the accumulators-only kernels
// do not depend on the store code.  Measured: C2 (EvenRosenbrock) gains 15 % from its store-free instantiation, C5 / C4 are
// unchanged (profiles/r1i_store_split_ab.txt).  RoughCarpet keeps ONE combined kernel with the round-1 store path: its
// store-free instantiation schedules 1.7 % worse than that (round 1: 5 %), and the combined kernel WITH the bulk-copy store
// code 3 % worse (profiles/r2_variant_ab.txt).  -DRWMPT_SPLIT_RC=1 builds the store-free RoughCarpet instantiation for A/B.
template <template <int, bool> class Target>
struct SplitStore {
  static constexpr bool value = true;
};
#ifndef RWMPT_SPLIT_RC
#define RWMPT_SPLIT_RC 0
#endif
template <>
struct SplitStore<RoughCarpet> {
  static constexpr bool value = RWMPT_SPLIT_RC != 0;
};
template <>
struct SplitStore<RoughCarpetPlain> {
  static constexpr bool value = RWMPT_SPLIT_RC != 0;
};

template <template <int, bool> class Target, int E, bool IEEE, int WT, int PF, bool EXACT, bool TEST, int VARIANT = 0>
cudaError_t launch_mcmc_one(const KernelArgs& a_in, const LaunchGeom& g, cudaStream_t st) {
  if constexpr (VARIANT >= 5) {
    // lean variants for CTAs of up to 64 threads, with or without retained samples
    if (g.threads > 64) return launch_mcmc_one<Target, E, IEEE, WT, PF, EXACT, TEST, 0>(a_in, g, st);
    if (a_in.samples != nullptr) return launch_mcmc_store<Target, E, IEEE, WT, PF, EXACT, TEST, true, VARIANT>(a_in, g, st);
    return launch_mcmc_store<Target, E, IEEE, WT, PF, EXACT, TEST, false, VARIANT>(a_in, g, st);
  } else if constexpr (VARIANT != 0) {
    // lean variants: accumulators-only runs of one-warp CTAs; anything else takes the pipelined kernel of the same geometry
    if (a_in.samples != nullptr || g.threads != 32) return launch_mcmc_one<Target, E, IEEE, WT, PF, EXACT, TEST, 0>(a_in, g, st);
    return launch_mcmc_store<Target, E, IEEE, WT, PF, EXACT, TEST, false, VARIANT>(a_in, g, st);
  } else if constexpr (TEST || !SplitStore<Target>::value) {
    return launch_mcmc_store<Target, E, IEEE, WT, PF, EXACT, TEST, true>(a_in, g, st);
  } else {
    if (a_in.samples != nullptr) return launch_mcmc_store<Target, E, IEEE, WT, PF, EXACT, TEST, true>(a_in, g, st);
    return launch_mcmc_store<Target, E, IEEE, WT, PF, EXACT, TEST, false>(a_in, g, st);
  }
}

template <template <int, bool> class Target, int E, bool IEEE, int WT, int PF, bool EXACT, bool TEST, bool STORE, int VARIANT>
cudaError_t launch_mcmc_store(const KernelArgs& a_in, const LaunchGeom& g, cudaStream_t st) {
  auto kern = mcmc_kernel<Target, E, IEEE, WT, PF, EXACT, TEST, STORE, VARIANT>;
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
  }
  // Balanced (time-sliced, ticketed) launch -- see mcmc_kernel.  Chosen when the plain grid would leave the SMs'
  // schedulers unevenly loaded in the few-warps-per-scheduler regime and every CTA of the SM-sized grid is resident.
  if (!TEST && g.schedule != RWMPT_SCHEDULE_PLAIN && g.sms > 0 && a_in.n_steps > 0 && g.grid < (1ll << 30) / 64) {
    const bool forced = g.schedule == RWMPT_SCHEDULE_BALANCED;
    const long long wpc = g.threads / 32;
    long long c = (g.grid + g.sms - 1) / g.sms;       // CTAs on the fullest SM of a plain launch
    while ((c * wpc) % 4) ++c;                         // same number of warps on each of the SM's four schedulers
    const long long P = (long long)g.sms * c;
    const long long min_slice = forced ? 16 : 4096;
    // measured on C3 (profiles/r1h_c3_schedule_vs_run_length.txt): the ticketed launch wins from ~1.5e5 steps per launch
    bool want = forced || (P > g.grid && c * wpc <= 16 && a_in.n_steps >= 40 * min_slice);
    if (want && a_in.n_steps >= 2 * min_slice) {
      int nb = 0;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, g.threads, g.smem);
      if (e != cudaSuccess) return e;
      if (nb >= c || forced) {
        KernelArgs a = a_in;
        long long S = a.n_steps / min_slice;
        if (S > 32) S = 32;
        long long len = (a.n_steps + S - 1) / S;
        len += len & 1;                                // slices start on a pair boundary of the Philox stream
        a.slice_steps = len;
        a.n_slices = (int)((a.n_steps + len - 1) / len);
        a.n_units = (int)g.grid;
        unsigned* ws = nullptr;
        const size_t bytes = ((size_t)g.grid + 1) * sizeof(unsigned);
        e = cudaMallocAsync((void**)&ws, bytes, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(ws, 0, bytes, st);
        if (e != cudaSuccess) {
          cudaFreeAsync(ws, st);
          return e;
        }
        a.ticket = ws;
        a.unit_done = reinterpret_cast<int*>(ws + 1);
        kern<<<(unsigned)P, g.threads, g.smem, st>>>(a);
        e = cudaGetLastError();
        cudaError_t e2 = cudaFreeAsync(ws, st);
        return e != cudaSuccess ? e : e2;
      }
    }
  }
  kern<<<(unsigned)g.grid, g.threads, g.smem, st>>>(a_in);
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
// decision outputs); fast kernels do not.  The two halves are separate function templates so that each family compiles as
// TWO translation units (rwmpt_inst_<family>.cu: fast, rwmpt_inst_<family>_ieee.cu: parity mode) -- 28 units instead of 14
// for the parallel build.
template <template <int, bool> class Target>
cudaError_t launch_mcmc_family_ieee(const KernelArgs& a, const LaunchGeom& g, cudaStream_t st) {
  switch (g.E) {
#define X(e) case e: return launch_mcmc_one<Target, e, true, 0, -1, false, true>(a, g, st);
    RWMPT_IEEE_E_LIST(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

template <template <int, bool> class Target>
cudaError_t launch_mcmc_family_fast(const KernelArgs& a, const LaunchGeom& g, cudaStream_t st) {
  cudaError_t e = Tuned<Target>::launch(a, g, st);
  if (e != cudaErrorNotSupported) return e;
  switch (g.E) {
#define X(e) case e: return launch_mcmc_one<Target, e, false, 0, -1, false, false>(a, g, st);
    RWMPT_FAST_E_LIST(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

template <template <int, bool> class Target>
cudaError_t launch_logp_family_ieee(const float* P, int d, int E, int W, const float* x, long long n, float* out, cudaStream_t st) {
  switch (E) {
#define X(e) case e: return launch_logp_one<Target, e, true>(P, d, W, x, n, out, st);
    RWMPT_IEEE_E_LIST(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

template <template <int, bool> class Target>
cudaError_t launch_logp_family_fast(const float* P, int d, int E, int W, const float* x, long long n, float* out, cudaStream_t st) {
  switch (E) {
#define X(e) case e: return launch_logp_one<Target, e, false>(P, d, W, x, n, out, st);
    RWMPT_FAST_E_LIST(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace rwmpt
#include "rwmpt_ladder.cuh"
namespace rwmpt {

// one set of entry points per family, defined in rwmpt_inst_<family>.cu
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
  X(mvn_diag, MVNDiag)                       \
  X(mvn_dense, MVNDense)                     \
  X(super_funnel, SuperFunnel)

#define X(name, cls)                                                                                         \
  cudaError_t launch_mcmc_##name(const KernelArgs& a, const LaunchGeom& g, bool ieee, cudaStream_t st);       \
  cudaError_t launch_logp_##name(const float* P, int d, int E, int W, const float* x, long long n, float* out, \
                                 bool ieee, cudaStream_t st);                                                 \
  cudaError_t launch_swap_prob_##name(const float* P, int d, int E, int W, float bc, float bs, long long n,     \
                                      unsigned k0, unsigned k1, long long row_base, double* sum_out, cudaStream_t st);
RWMPT_FAMILY_LIST(X)
#undef X

// RWMPT_DEFINE_TUNED(cls, LIST) with LIST(X) = X(E, W, PF) ... specialises Tuned<cls>
#define RWMPT_TUNED_CASE(cls, e, w, pf)                                                        \
  if (g.E == e && g.W == w && a.prop_family == pf) {                              \
    if (a.dim == e * w) return launch_mcmc_one<cls, e, false, w, pf, true, false>(a, g, st);  \
    return launch_mcmc_one<cls, e, false, w, pf, false, false>(a, g, st);                     \
  }
// lean-loop variant v (1 or 2, see mcmc_kernel) of a tuned case; list it BEFORE the pipelined case of the same geometry
#define RWMPT_TUNED_CASE_V(cls, e, w, pf, v)                                                         \
  if (g.variant == v && g.E == e && g.W == w && a.prop_family == pf) {                              \
    if (a.dim == e * w) return launch_mcmc_one<cls, e, false, w, pf, true, false, v>(a, g, st);     \
    return launch_mcmc_one<cls, e, false, w, pf, false, false, v>(a, g, st);                        \
  }
// tuned case that swaps in a leaner functor (e.g. RoughCarpetPlain) when the target carries no scaling block
#define RWMPT_TUNED_PLAIN_CASE(cls, plain, e, w, pf)                                               \
  if (a.target_plain && g.E == e && g.W == w && a.prop_family == pf) {                             \
    if (a.dim == e * w) return launch_mcmc_one<plain, e, false, w, pf, true, false>(a, g, st);    \
    return launch_mcmc_one<plain, e, false, w, pf, false, false>(a, g, st);                       \
  }
#define RWMPT_TUNED_PLAIN_CASE_V(cls, plain, e, w, pf, v)                                               \
  if (g.variant == v && a.target_plain && g.E == e && g.W == w && a.prop_family == pf) {              \
    if (a.dim == e * w) return launch_mcmc_one<plain, e, false, w, pf, true, false, v>(a, g, st);     \
    return launch_mcmc_one<plain, e, false, w, pf, false, false, v>(a, g, st);                        \
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
  cudaError_t launch_mcmc_ieee_##name(const KernelArgs& a, const LaunchGeom& g, cudaStream_t st);             \
  cudaError_t launch_logp_ieee_##name(const float* P, int d, int E, int W, const float* x, long long n,       \
                                      float* out, cudaStream_t st);                                           \
  cudaError_t launch_mcmc_##name(const KernelArgs& a, const LaunchGeom& g, bool ieee, cudaStream_t st) {      \
    return ieee ? launch_mcmc_ieee_##name(a, g, st) : launch_mcmc_family_fast<cls>(a, g, st);                 \
  }                                                                                                           \
  cudaError_t launch_logp_##name(const float* P, int d, int E, int W, const float* x, long long n, float* out, \
                                 bool ieee, cudaStream_t st) {                                                \
    return ieee ? launch_logp_ieee_##name(P, d, E, W, x, n, out, st)                                          \
                : launch_logp_family_fast<cls>(P, d, E, W, x, n, out, st);                                    \
  }                                                                                                           \
  cudaError_t launch_swap_prob_##name(const float* P, int d, int E, int W, float bc, float bs, long long n,    \
                                      unsigned k0, unsigned k1, long long row_base, double* sum_out,          \
                                      cudaStream_t st) {                                                      \
    return launch_swap_prob_family<cls>(P, d, E, W, bc, bs, n, k0, k1, row_base, sum_out, st);                \
  }                                                                                                           \
  }

// the parity-mode half of a family, in its own translation unit (rwmpt_inst_<family>_ieee.cu)
#define RWMPT_DEFINE_FAMILY_IEEE(name, cls)                                                                   \
  namespace rwmpt {                                                                                           \
  cudaError_t launch_mcmc_ieee_##name(const KernelArgs& a, const LaunchGeom& g, cudaStream_t st) {            \
    return launch_mcmc_family_ieee<cls>(a, g, st);                                                            \
  }                                                                                                           \
  cudaError_t launch_logp_ieee_##name(const float* P, int d, int E, int W, const float* x, long long n,       \
                                      float* out, cudaStream_t st) {                                          \
    return launch_logp_family_ieee<cls>(P, d, E, W, x, n, out, st);                                           \
  }                                                                                                           \
  }

}  // namespace rwmpt
