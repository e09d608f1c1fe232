// rwmpt_common.cuh -- shared device utilities for the sm_100a RWM / PT-RWM kernels:
// kernel argument block, Philox4x32-10, math policies (fast intrinsics vs IEEE parity), sub-warp reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rwmpt.h"

namespace rwmpt {

constexpr int kMaxCtaThreads = 256;  // a whole ladder (n_temps * lanes_per_chain threads) must fit one CTA

// Device-side argument block (passed by value as the single kernel parameter).
struct KernelArgs {
  const float* P;  // target parameters (device)
  int dim;
  int prop_family;
  int K;  // temperatures per ladder
  int W;  // lanes per chain (power of two <= 32)
  int chains_per_cta;
  int swap_every;
  int swap_mode;
  int store_mode;
  const float* prop_scale;
  const float* prop_dim_scale;
  const float* beta;
  long long n_chains;
  long long n_steps;
  long long burn_in;
  long long step_offset;
  long long rounds_before;  // swap sweeps performed before step_offset (global round numbering)
  float* state;
  float* logp;
  unsigned int key0, key1;
  unsigned int rk[20];  // Philox round keys (key0 + r*W0, key1 + r*W1), r = 0..9
  long long chain_id_base;
  float* samples;
  float* sample_logp;
  long long store_start, thin, sample_stride, sample_rows;
  unsigned long long* accept_count;
  double* sq_jump_sum;
  unsigned long long* swap_accepts;
  unsigned long long* swap_last_attempt;
  const float* inj_inc;
  const float* inj_u;
  const float* inj_su;
  unsigned char* decisions;
  unsigned char* swap_dec;
  long long n_ladders;
  int target_plain;  // the target has no per-coordinate scaling block (RoughCarpet: n_params == header)
  int stage_rows;  // rows per chain staged in shared memory before a coalesced flush
  int stage_off;   // offset (floats) of the staging region in dynamic shared memory
  int stage_vw;    // floats per vector store of the flush (4, 2 or 1)
  int stage_bufs;  // 2: double-buffered staging, blocks leave through the bulk-copy engine (cp.async.bulk shared -> global)
  int stage_stride;  // floats between two staging buffers: stage_rows * dim plus the slack the unmasked row stores spill into
  // balanced (time-sliced, ticketed) launch -- see mcmc_kernel; n_slices <= 1: plain launch, CTA b runs unit b
  int n_slices;
  int n_units;            // CTAs of the plain launch (= units of work)
  long long slice_steps;  // steps per slice (even)
  unsigned* ticket;       // [1] zero-initialised
  int* unit_done;         // [n_units] zero-initialised: slices published per unit
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) -- counter-based, no state in memory.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                      uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  // one IMAD.WIDE.U32 per product (4 issue cycles on B200) instead of IMAD.HI (4) + IMAD (2): measured with
  // scripts/probes/imad_probe.cu.  The asm keeps ptxas from splitting the 64-bit product.
  unsigned long long p0, p1;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(c0), "r"(M0));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(c2), "r"(M1));
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#else
  const uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c0) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c2) >> 32);
  const uint32_t lo0 = M0 * c0, lo1 = M1 * c2;
#endif
  const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
  c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += W0;
    k1 += W1;
  }
  uint4 o;
  o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// torch.rand semantics: 24 random bits -> [0, 1)
__device__ __forceinline__ float u01_from_bits(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
// (0, 1]: never 0, so log() is finite
__device__ __forceinline__ float u01_open_low(uint32_t w) { return fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// ------------------------------------------------------------------------------------------------
// Math policies.  IEEE = parity mode: every operation individually rounded (no FMA contraction),
// accurate expf/logf, in the reference's operation order.  Fast = MUFU approximations, FMA allowed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell packed fp32 (FFMA2 / FADD2 / FMUL2): two fp32 operations per issued instruction.  The fused kernel is
// bound by instruction issue, not by the fp32 pipe, so the density loops pack adjacent coordinates.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ---- bulk asynchronous copy shared::cta -> global (the TMA engine's 1-D form): one thread hands a contiguous, 16-byte
// aligned block to the copy engine; the SM's load/store pipes never see the data again.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until all but the N most recent bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// make this thread's generic-proxy writes to shared memory visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <bool IEEE>
struct Mth;

template <>
struct Mth<true> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sq(float a) { return __fmul_rn(a, a); }
  static __device__ __forceinline__ float exp(float a) { return expf(a); }
  static __device__ __forceinline__ float log(float a) { return logf(a); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};

template <>
struct Mth<false> {
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float div(float a, float b) { return a * rcp_approx(b); }
  static __device__ __forceinline__ float sq(float a) { return a * a; }
  static __device__ __forceinline__ float exp(float a) { return ex2_approx(a * kLog2e); }
  static __device__ __forceinline__ float log(float a) { return lg2_approx(a) * kLn2; }
  static __device__ __forceinline__ float sqrt(float a) { return sqrt_approx(a); }
};

// ------------------------------------------------------------------------------------------------
// Sub-warp reductions: the W lanes of a chain are W consecutive lanes of one warp (W power of two).
// XOR butterflies: every lane of the group ends with the bit-identical result, so all lanes of a chain
// take the same accept / swap decision without a broadcast.
// ------------------------------------------------------------------------------------------------
constexpr unsigned kFull = 0xffffffffu;

// WT > 0: lanes-per-chain known at compile time (fully unrolled butterfly); WT == 0: runtime W.
template <int WT>
__device__ __forceinline__ float group_sum_w(float v, int W) {
  if constexpr (WT > 0) {
#pragma unroll
    for (int o = WT / 2; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, o));
  } else {
    if (W > 16) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, 16));
    if (W > 8) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, 8));
    if (W > 4) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, 4));
    if (W > 2) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, 2));
    if (W > 1) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, 1));
  }
  return v;
}

template <int WT>
__device__ __forceinline__ double group_sum_f64_w(double v, int W) {
  if constexpr (WT > 0) {
#pragma unroll
    for (int o = WT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  } else {
    if (W > 16) v += __shfl_xor_sync(kFull, v, 16);
    if (W > 8) v += __shfl_xor_sync(kFull, v, 8);
    if (W > 4) v += __shfl_xor_sync(kFull, v, 4);
    if (W > 2) v += __shfl_xor_sync(kFull, v, 2);
    if (W > 1) v += __shfl_xor_sync(kFull, v, 1);
  }
  return v;
}

// Per-thread view of where it sits inside its chain.
template <int WT_, bool EXACT_>
struct CtxT {
  static constexpr int WT = WT_;
  static constexpr bool EXACT = EXACT_;  // E * W == d: no padding coordinates, masks fold away
  __device__ __forceinline__ bool ok(int e) const {
    if constexpr (EXACT_) return true;
    else return base + e < d;
  }
  const float* P;  // target params
  int d;           // dimension
  int W;           // lanes per chain
  int sub;         // this lane's index inside the chain group [0, W)
  int base;        // first global coordinate held by this lane (= sub * E)
  int lane;        // lane in warp
  int leader;      // lane (in warp) of sub == 0 of this chain
};
using Ctx = CtxT<0, false>;

template <class C>
__device__ __forceinline__ float group_sum(float v, const C& c) { return group_sum_w<C::WT>(v, c.W); }

// value of `v` held by lane `sub+delta` of the same chain (garbage at the group edge: callers mask it)
__device__ __forceinline__ float from_next_lane(float v) { return __shfl_down_sync(kFull, v, 1); }
__device__ __forceinline__ float from_prev_lane(float v) { return __shfl_up_sync(kFull, v, 1); }
template <class C>
__device__ __forceinline__ float from_leader(float v, const C& c) { return __shfl_sync(kFull, v, c.leader); }

}  // namespace rwmpt
