// rwmpt_ladder.cuh -- native swap-probability estimator behind the iterative temperature-ladder construction
// (reference: ParallelTemperingRWM_GPU_Optimized._construct_iterative_ladder, algorithms/pt_rwm_gpu_optimized.py:283-426;
// the estimate itself is :356-368).
//
// One launch per (beta, beta*) pair:  a = mean_n min(1, exp((beta - beta*) (log pi(x*_n) - log pi(x_n)))),  x_n drawn by the
// target's heuristic tempered sampler at beta and x*_n at beta* (`draw_samples_torch` of target_distributions/*_torch.py:
// multimodal_torch.py:270-333, 532-565; rosenbrock_torch.py:224-248; multivariate_normal_torch.py:101-121, 249-268).
// Everything the reference does with five tensor passes over (N, d) arrays -- two sample draws, two log-densities, the
// clamped exponential and the mean -- happens in registers: in-kernel Philox, the same density functors as the sampling
// kernel, one fp64 atomic per CTA.  HBM traffic: the target parameters in, 8 bytes out.
#pragma once

#include "rwmpt_kernel.cuh"

namespace rwmpt {

// Randomness of one sample row: Philox counter = (coordinate pair index, stream tag, row id); one call yields the two
// normals (Box-Muller) and the two uniforms of coordinates (2b, 2b+1).  Counter-based, so a coordinate that needs its
// neighbour's draw (EvenRosenbrock: x_{2j+1} | x_{2j}) simply regenerates it.
struct RowRng {
  unsigned k0, k1;
  unsigned long long rid;
  unsigned tag;
  int cached = -1;
  float z0, z1, u0, u1;
  __device__ __forceinline__ void load(int block) {
    if (block == cached) return;
    const uint4 w = philox4x32_10((unsigned)block, tag, (unsigned)rid, (unsigned)(rid >> 32), k0, k1);
    box_muller<false>(w.x, w.y, z0, z1);
    u0 = u01_from_bits(w.z);
    u1 = u01_from_bits(w.w);
    cached = block;
  }
  __device__ __forceinline__ float normal(int i) { load(i >> 1); return (i & 1) ? z1 : z0; }
  __device__ __forceinline__ float uniform(int i) { load(i >> 1); return (i & 1) ? u1 : u0; }
  // one uniform per row (mixture component of ThreeMixture): a block index no coordinate uses
  __device__ __forceinline__ float row_uniform() const {
    const uint4 w = philox4x32_10(0xffffffffu, tag, (unsigned)rid, (unsigned)(rid >> 32), k0, k1);
    return u01_from_bits(w.x);
  }
};

__device__ __forceinline__ int pick3(float u, float w0, float w1) { return u < w0 ? 0 : (u < w0 + w1 ? 1 : 2); }

// Tempered<Target>::coord(P, d, i, beta, rng): coordinate i of one heuristic sample at inverse temperature beta.
template <template <int, bool> class Target>
struct Tempered {
  static constexpr bool supported = false;
  struct Row { __device__ void init(const float*, int, float, RowRng&) {} };
  static __device__ float coord(const float*, int, int, float, const Row&, RowRng&) { return 0.0f; }
};

// RoughCarpet (multimodal_torch.py:532-565): per coordinate, mode ~ weights, x_i = (m + z / sqrt(beta)) / s_i
template <>
struct Tempered<RoughCarpet> {
  static constexpr bool supported = true;
  struct Row {
    float w0, w1, isb;
    __device__ __forceinline__ void init(const float* P, int, float beta, RowRng&) {
      w0 = __expf(P[3]); w1 = __expf(P[4]); isb = rsqrtf(beta);
    }
  };
  static __device__ __forceinline__ float coord(const float* P, int, int i, float, const Row& r, RowRng& g) {
    const int k = pick3(g.uniform(i), r.w0, r.w1);
    const float y = P[k] + g.normal(i) * r.isb;
    return P[7] != 0.0f ? y / P[RWMPT_PARAM_HEADER + i] : y;
  }
};

// ThreeMixture (multimodal_torch.py:270-333): one component per ROW, x = (mu_k + z / sqrt(beta)) / s
template <>
struct Tempered<ThreeMixture> {
  static constexpr bool supported = true;
  struct Row {
    int k; float isb;
    __device__ __forceinline__ void init(const float* P, int, float beta, RowRng& g) {
      k = pick3(g.row_uniform(), __expf(P[0]), __expf(P[1]));
      isb = rsqrtf(beta);
    }
  };
  static __device__ __forceinline__ float coord(const float* P, int d, int i, float, const Row& r, RowRng& g) {
    const float y = P[RWMPT_PARAM_HEADER + r.k * d + i] + g.normal(i) * r.isb;
    return P[6] != 0.0f ? y / P[RWMPT_PARAM_HEADER + 3 * d + i] : y;
  }
};

// EvenRosenbrock (rosenbrock_torch.py:224-248): x_{2j} ~ N(mu_j, 1 / (2 a beta)), x_{2j+1} | x_{2j} ~ N(x_{2j}^2, 1 / (2 b beta))
template <>
struct Tempered<EvenRosenbrock> {
  static constexpr bool supported = true;
  struct Row {
    float sa, sb;
    __device__ __forceinline__ void init(const float* P, int, float beta, RowRng&) {
      const float ea = P[0] * beta, eb = P[1] * beta;
      sa = ea > 0.0f ? sqrtf(1.0f / (2.0f * ea)) : 1.0f;
      sb = eb > 0.0f ? sqrtf(1.0f / (2.0f * eb)) : 1.0f;
    }
  };
  static __device__ __forceinline__ float coord(const float* P, int, int i, float, const Row& r, RowRng& g) {
    const float even = P[RWMPT_PARAM_HEADER + (i >> 1)] + g.normal(i & ~1) * r.sa;
    if ((i & 1) == 0) return even;
    return even * even + g.normal(i) * r.sb;
  }
};

// ScaledMultivariateNormal (multivariate_normal_torch.py:249-268): x_i ~ N(0, 1 / (c_i^2 beta))
template <>
struct Tempered<ScaledMVN> {
  static constexpr bool supported = true;
  struct Row {
    float isb;
    __device__ __forceinline__ void init(const float*, int, float beta, RowRng&) { isb = rsqrtf(beta); }
  };
  static __device__ __forceinline__ float coord(const float* P, int, int i, float, const Row& r, RowRng& g) {
    return g.normal(i) * r.isb / P[RWMPT_PARAM_HEADER + i];
  }
};

// MultivariateNormal with diagonal covariance (multivariate_normal_torch.py:101-121): x = mean + chol(cov / beta) z
template <>
struct Tempered<MVNDiag> {
  static constexpr bool supported = true;
  struct Row {
    float isb;
    __device__ __forceinline__ void init(const float*, int, float beta, RowRng&) { isb = rsqrtf(beta); }
  };
  static __device__ __forceinline__ float coord(const float* P, int d, int i, float, const Row& r, RowRng& g) {
    return P[RWMPT_PARAM_HEADER + i] + g.normal(i) * r.isb * rsqrtf(P[RWMPT_PARAM_HEADER + d + i]);
  }
};

// Same lane mapping as logp_kernel: W lanes per row, E coordinates per lane; grid-stride over rows.
template <template <int, bool> class Target, int E>
__global__ void __launch_bounds__(128) swap_prob_kernel(const float* __restrict__ P, int d, int W, float beta_curr, float beta_star,
                                                        long long n, unsigned k0, unsigned k1, long long row_base,
                                                        double* __restrict__ sum_out) {
  using T = Tempered<Target>;
  const int per_cta = blockDim.x / W;
  const int cl = threadIdx.x / W;
  CtxT<0, false> c;
  c.P = P; c.d = d; c.W = W;
  c.sub = threadIdx.x % W;
  c.base = c.sub * E;
  c.lane = threadIdx.x & 31;
  c.leader = c.lane & ~(W - 1);
  Target<E, false> tgt;
  tgt.init(c);
  const long long stride = (long long)gridDim.x * per_cta;
  const long long n_round = ((n + stride - 1) / stride) * stride;  // keep warps converged for the shuffles
  float acc = 0.0f;
  for (long long r = (long long)blockIdx.x * per_cta + cl; r < n_round; r += stride) {
    const bool ok = r < n;
    const unsigned long long rid = (unsigned long long)(row_base + (ok ? r : 0));
    float lp[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {  // s = 0: sample at beta_curr, s = 1: sample at beta_star (independent streams)
      RowRng g{k0, k1, rid, 0x4C414444u + (unsigned)s};
      const float beta = s ? beta_star : beta_curr;
      typename T::Row row;
      row.init(P, d, beta, g);
      float x[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = c.base + e;
        x[e] = i < d ? T::coord(P, d, i, beta, row, g) : 0.0f;
      }
      lp[s] = tgt.logp(x, c);
    }
    const float log_r = (beta_curr - beta_star) * (lp[1] - lp[0]);          // :364
    const float a = __expf(fminf(log_r, 0.0f));                              // exp(clamp_max(., 0)) = min(1, exp(.))  :366
    acc += (ok && c.sub == 0 && a == a) ? a : 0.0f;
  }
  double tot = (double)acc;
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
  __shared__ double s_acc[4];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(sum_out, (s_acc[0] + s_acc[1]) + (s_acc[2] + s_acc[3]));
}

template <template <int, bool> class Target, int E>
cudaError_t launch_swap_prob_one(const float* P, int d, int W, float bc, float bs, long long n, unsigned k0, unsigned k1,
                                 long long row_base, double* sum_out, cudaStream_t st) {
  if constexpr (!Tempered<Target>::supported) {
    return cudaErrorNotSupported;
  } else {
    const int per_cta = 128 / W;
    long long blocks = (n + per_cta - 1) / per_cta;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    swap_prob_kernel<Target, E><<<(unsigned)blocks, 128, 0, st>>>(P, d, W, bc, bs, n, k0, k1, row_base, sum_out);
    return cudaGetLastError();
  }
}

template <template <int, bool> class Target>
cudaError_t launch_swap_prob_family(const float* P, int d, int E, int W, float bc, float bs, long long n, unsigned k0, unsigned k1,
                                    long long row_base, double* sum_out, cudaStream_t st) {
  if constexpr (!Tempered<Target>::supported) {
    return cudaErrorNotSupported;
  } else {
    switch (E) {
#define X(e) case e: return launch_swap_prob_one<Target, e>(P, d, W, bc, bs, n, k0, k1, row_base, sum_out, st);
      RWMPT_FAST_E_LIST(X)
#undef X
    }
    return cudaErrorInvalidValue;
  }
}

}  // namespace rwmpt
