// rwmpt_kernel.cuh -- the persistent fused multi-step RWM / PT-RWM kernel for sm_100a.
//
// One launch runs all n_steps Metropolis steps of every chain.  Chain state, log-density, proposal
// scale, beta, target parameters and all accumulators live in registers for the whole run; per step a
// chain draws its increments and its accept-uniform from in-kernel Philox4x32-10, evaluates the target
// functor on the proposal, reduces over the chain's lanes with XOR shuffles and accepts / rejects.
// A PT ladder (n_temps chains) is resident in one CTA; every swap_every steps the CTA runs the
// adjacent-temperature sweep through shared memory.  HBM sees only the initial load, the final store,
// and (optionally) retained samples / decisions.
//
// Thread mapping: a chain's d coordinates are split over W consecutive lanes of a warp, E coordinates per
// lane held in registers (blocked).  CTA = chains_per_cta chains = whole ladders; small CTAs (usually one
// warp) so that B chains spread evenly over the 148 SMs.
//
// Replaces: _single_step_ultra_fused + ultra_fused_mcmc_step_basic (rwm_gpu_optimized.py:289-336, 9-32),
// step + ultra_fused_parallel_mcmc_step + _attempt_all_swaps + _add_states_to_chains
// (pt_rwm_gpu_optimized.py:541-574, 61-84, 594-633, 635-653), and the proposal plugins' samplers.
#pragma once

#include <type_traits>

#include "rwmpt_common.cuh"
#include "rwmpt_targets.cuh"

#ifndef RWMPT_CHUNK
#define RWMPT_CHUNK 32  // pairs of steps between two flushes of the fp32 partial sums into the fp64 / 64-bit accumulators
#endif
#ifndef RWMPT_ORDER
#define RWMPT_ORDER 0  // source order of the three streams inside the fast loop (ptxas keeps it as a tie-break)
#endif

namespace rwmpt {

#ifdef RWMPT_NO_F32X2
constexpr bool kUseF32x2 = false;
#else
constexpr bool kUseF32x2 = true;  // packed fp32 (FFMA2 / FADD2 / FMUL2) in the fast-math paths
#endif

// ---- Philox with the 10 round keys precomputed on the host (uniform operands from the argument block) -----
__device__ __forceinline__ uint4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const KernelArgs& a) {
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, a.rk[2 * r], a.rk[2 * r + 1]);
  uint4 o;
  o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// ---- normal pair from two Philox words; `rs2` = -2 ln2 * scale^2 folds the proposal std into the radius ----
template <bool IEEE>
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1, float rs2 = -2.0f * kLn2) {
  const float u1 = u01_open_low(wa);
  if constexpr (IEEE) {
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif((float)wb * 4.6566128730773926e-10f, &sn, &cs);  // angle = 2*pi*wb/2^32
    z0 = r * cs;
    z1 = r * sn;
  } else {
    const float r = sqrt_approx(rs2 * lg2_approx(u1));
    const float th = (float)wb * 1.4629180792671596e-9f;  // 2*pi/2^32
    z0 = r * __cosf(th);
    z1 = r * __sinf(th);
  }
}

// Words one lane needs for TWO consecutive steps: 2E normals/uniforms + 2 accept uniforms (+ 2 ball radii).
__host__ __device__ constexpr int pair_words(int E, int PF) {
  return 2 * E + 2 + ((PF == RWMPT_P_NORMAL || PF == RWMPT_P_LAPLACE) ? 0 : 2);
}

// Randomness of one lane for the two steps of pair `pair` (global steps 2*pair+1 and 2*pair+2): increments after
// scaling, and the chain's accept uniforms (taken from the leader lane).  Drawing two steps at once uses every
// Philox word (E=5, Normal: 3 calls per 2 steps) and gives the scheduler three independent Philox chains.
// PF >= 0: proposal family known at compile time; PF < 0: runtime switch on a.prop_family.
//
// Counter layout of call k of lane `sub`:  c0 = (pair >> 32) << 16 | sub << 8 | k,  c1 = (uint32) pair,
// (c2, c3) = global chain id; key = seed.  Only c1 -- an XOR input of the first round -- changes from pair to pair, so
// everything that does not depend on it is evaluated once per run (PhiloxPairGen): round 1 costs one XOR shared by the
// lane's calls, round 2 one multiply (shared) and one XOR, round 3 one multiply and two XORs; rounds 4..10 are plain.
// The words are bit-identical to philox4x32_10 on that counter (PairWords / tests/test_gpu_parity.py resume tests).
__host__ __device__ __forceinline__ void mulhilo32(uint32_t m, uint32_t x, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  unsigned long long p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x), "r"(m));
  hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
#else
  const uint64_t p = (uint64_t)m * x;
  hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ uint32_t pair_c0(unsigned long long pair, int sub, int k) {
  return ((uint32_t)(pair >> 32) << 16) | ((uint32_t)sub << 8) | (uint32_t)k;
}

template <int E, int PF>
struct PairWords {
  static constexpr int NW = pair_words(E, PF);
  static constexpr int NC = (NW + 3) / 4;
  uint32_t w[4 * NC];
  // reference evaluation: ten plain rounds per call
  __host__ __device__ __forceinline__ void full(const uint32_t* rk, int sub, unsigned long long pair, unsigned long long chain_gid) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      uint32_t c0 = pair_c0(pair, sub, k), c1 = (uint32_t)pair, c2 = (uint32_t)chain_gid, c3 = (uint32_t)(chain_gid >> 32);
#pragma unroll
      for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, rk[2 * r], rk[2 * r + 1]);
      w[4 * k] = c0; w[4 * k + 1] = c1; w[4 * k + 2] = c2; w[4 * k + 3] = c3;
    }
  }
};

template <int NC>
struct PhiloxPairGen {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t A1;                            // round 1: n0 = c1 ^ A1 (same for every call of the lane)
  uint32_t B2[NC], C3[NC], D3[NC], E4[NC];  // per-call invariants of rounds 2..4
  uint32_t hi32;                          // pair >> 32 the invariants were built for
  __host__ __device__ __forceinline__ void init(const uint32_t* rk, int sub, unsigned long long pair, unsigned long long chain_gid) {
    hi32 = (uint32_t)(pair >> 32);
    const uint32_t c2 = (uint32_t)chain_gid, c3 = (uint32_t)(chain_gid >> 32);
    uint32_t h1, l1;
    mulhilo32(M1, c2, h1, l1);
    A1 = h1 ^ rk[0];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      uint32_t h0, l0, h1p, l1p, h0pp, l0pp;
      mulhilo32(M0, pair_c0(pair, sub, k), h0, l0);
      const uint32_t N2 = h0 ^ c3 ^ rk[1];         // round 1 -> (c1 ^ A1, l1, N2, l0)
      mulhilo32(M1, N2, h1p, l1p);
      const uint32_t N0p = h1p ^ l1 ^ rk[2];       // round 2 -> (N0p, l1p, hi(M0 n0) ^ B2, lo(M0 n0))
      B2[k] = l0 ^ rk[3];
      mulhilo32(M0, N0p, h0pp, l0pp);              // round 3 -> (hi(M1 n2') ^ C3, lo(M1 n2'), lo(M0 n0) ^ D3, l0pp)
      C3[k] = l1p ^ rk[4];
      D3[k] = h0pp ^ rk[5];
      E4[k] = l0pp ^ rk[7];                        // round 4: n2 = hi(M0 x0) ^ E4
    }
  }
  __host__ __device__ __forceinline__ void gen(const uint32_t* rk, uint32_t pair_lo, uint32_t (&w)[4 * NC]) const {
    uint32_t hA, lA;
    mulhilo32(M0, pair_lo ^ A1, hA, lA);           // rounds 1 and 2, shared by the lane's calls
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      uint32_t h1, l1v, h0, l0n, h1b, l1n;
      mulhilo32(M1, hA ^ B2[k], h1, l1v);          // round 3
      const uint32_t x0 = h1 ^ C3[k], x2 = lA ^ D3[k];
      mulhilo32(M0, x0, h0, l0n);                  // round 4
      mulhilo32(M1, x2, h1b, l1n);
      uint32_t c0 = h1b ^ l1v ^ rk[6], c1 = l1n, c2 = h0 ^ E4[k], c3 = l0n;
#pragma unroll
      for (int r = 4; r < 10; ++r) philox_round(c0, c1, c2, c3, rk[2 * r], rk[2 * r + 1]);
      w[4 * k] = c0; w[4 * k + 1] = c1; w[4 * k + 2] = c2; w[4 * k + 3] = c3;
    }
  }
};

template <int E, bool IEEE, int PF, class C>
__device__ __forceinline__ void pair_transform(const KernelArgs& a, const C& c, const uint32_t (&w)[4 * PairWords<E, PF>::NC],
                                               float (&incA)[E], float (&incB)[E], float& uA, float& uB, uint32_t& sA,
                                               uint32_t& sB, float scale, const float (&dscale)[E]) {
  constexpr int NC = PairWords<E, PF>::NC;
  const int pf = PF >= 0 ? PF : a.prop_family;
  uA = from_leader(u01_from_bits(w[4 * NC - 1]), c);
  uB = from_leader(u01_from_bits(w[4 * NC - 2]), c);
  // Only the leader lane's two accept words are consumed; the same words of the chain's second lane feed the swap sweep
  // that may follow step A / step B (see `sweep`), so a sweep costs no Philox call of its own.
  sA = w[4 * NC - 1];
  sB = w[4 * NC - 2];
  if constexpr (!IEEE) {  // fast path compares ln(u) < lar (see mh_accept)
    uA = lg2_approx(uA) * kLn2;
    uB = lg2_approx(uB) * kLn2;
  }
  if constexpr (!IEEE) {
    if (pf == RWMPT_P_LAPLACE) {
      // laplace.py:47-69 without a conversion or a branch: 23 word bits -> f in [1, 2); r = 2f - 3 = 2u in [-1, 1);
      // log1p(max(-|r|, -0.999999)) = ln(max(1 - |r|, 1e-6)) <= 0, so the increment is |s ln(.)| with the sign of u.
      // `dscale` is 0 on padding coordinates (mcmc_unit), which keeps their increment at 0.
      float sdl[E];
#pragma unroll
      for (int e = 0; e < E; ++e) sdl[e] = scale * dscale[e] * kLn2;
#pragma unroll
      for (int e = 0; e < 2 * E; ++e) {
        const float f = __uint_as_float((w[e] & 0x007fffffu) | 0x3f800000u);
        const float r = fmaf(2.0f, f, -3.0f);
        const float m = lg2_approx(fmaxf(1.0f - fabsf(r), 1e-6f)) * sdl[e < E ? e : e - E];
        const float v = copysignf(m, r);
        if (e < E) incA[e] = v; else incB[e - E] = v;
      }
      return;
    }
  }
  if (pf == RWMPT_P_LAPLACE) {
    // laplace.py:47-69: u in (-.5,.5); -s * sign(u) * log1p(max(-2|u|, -0.999999))
#pragma unroll
    for (int e = 0; e < 2 * E; ++e) {
      const float u = u01_from_bits(w[e]) - 0.5f;
      const float arg = fmaxf(-2.0f * fabsf(u), -0.999999f);
      const float l = IEEE ? log1pf(arg) : lg2_approx(1.0f + arg) * kLn2;
      const float sg = (u > 0.0f) ? 1.0f : ((u < 0.0f) ? -1.0f : 0.0f);
      const float v = -(scale * dscale[e < E ? e : e - E]) * sg * l;
      if (e < E) incA[e] = v; else incB[e - E] = v;
    }
    return;
  }
#ifndef RWMPT_NO_F32X2
  if constexpr (!IEEE && PF == RWMPT_P_NORMAL && E >= 2) {
    // Packed-fp32 Box-Muller: two (radius, angle) pairs per FFMA2 / FMUL2.  The 2E normals of the lane are iid, so they
    // are dealt to the coordinates of the two steps in the order that leaves (e, e+1) in one register pair.
    constexpr int H = E / 2;
    const float rs2s = -2.0f * kLn2 * scale * scale;
    const f32x2_t RS2 = pack2(rs2s, rs2s);
    const f32x2_t KU = pack2(2.3283064365386963e-10f, 2.3283064365386963e-10f);
    const f32x2_t KU0 = pack2(1.1641532182693481e-10f, 1.1641532182693481e-10f);
    const f32x2_t KT = pack2(1.4629180792671596e-9f, 1.4629180792671596e-9f);
#pragma unroll
    for (int g = 0; g < H; ++g) {
      float ua, ub, ta, tb, qa, qb;
      unpack2(fma2(pack2((float)w[4 * g], (float)w[4 * g + 2]), KU, KU0), ua, ub);
      unpack2(mul2(pack2((float)w[4 * g + 1], (float)w[4 * g + 3]), KT), ta, tb);
      unpack2(mul2(RS2, pack2(lg2_approx(ua), lg2_approx(ub))), qa, qb);
      const f32x2_t r = pack2(sqrt_approx(qa), sqrt_approx(qb));
      float z0, z1, z2, z3;
      unpack2(mul2(r, pack2(__cosf(ta), __cosf(tb))), z0, z1);
      unpack2(mul2(r, pack2(__sinf(ta), __sinf(tb))), z2, z3);
      // pair slots 2g and 2g+1: the first H slots belong to step A, the rest to step B
      if (2 * g < H) { incA[4 * g] = z0; incA[4 * g + 1] = z1; } else { incB[4 * g - 2 * H] = z0; incB[4 * g - 2 * H + 1] = z1; }
      if (2 * g + 1 < H) { incA[4 * g + 2] = z2; incA[4 * g + 3] = z3; } else { incB[4 * g + 2 - 2 * H] = z2; incB[4 * g + 3 - 2 * H] = z3; }
    }
    if constexpr (E & 1) box_muller<false>(w[4 * H], w[4 * H + 1], incA[E - 1], incB[E - 1], rs2s);
    return;
  }
#endif
  float z[2 * E];
  const bool fold = !IEEE && pf == RWMPT_P_NORMAL;  // normal.py:47-55: randn * std
  const float rs2 = fold ? -2.0f * kLn2 * scale * scale : -2.0f * kLn2;
#pragma unroll
  for (int p = 0; p < E; ++p) box_muller<IEEE>(w[2 * p], w[2 * p + 1], z[2 * p], z[2 * p + 1], rs2);
  if (pf == RWMPT_P_NORMAL) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      incA[e] = fold ? z[e] : z[e] * scale;
      incB[e] = fold ? z[E + e] : z[E + e] * scale;
    }
  } else {
    // uniform.py:48-73: z/||z|| * R * u^(1/d)
    float nA = 0.0f, nB = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (c.ok(e)) { nA = fmaf(z[e], z[e], nA); nB = fmaf(z[E + e], z[E + e], nB); }
    nA = group_sum(nA, c);
    nB = group_sum(nB, c);
    const float inv_d = 1.0f / (float)c.d;
    float f[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float n2 = h ? nB : nA;
      const float nrm = IEEE ? sqrtf(n2) : sqrt_approx(n2);
      const float safe = nrm > 1e-12f ? nrm : 1.0f;
      const float ur = from_leader(u01_from_bits(w[4 * NC - 3 - h]), c);
      const float rad = IEEE ? scale * powf(ur, inv_d) : scale * ex2_approx(lg2_approx(ur) * inv_d);
      f[h] = IEEE ? rad / safe : rad * rcp_approx(safe);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) { incA[e] = z[e] * f[0]; incB[e] = z[E + e] * f[1]; }
  }
}

template <int E, bool IEEE, int PF, class C>
__device__ __forceinline__ void draw_pair(const KernelArgs& a, const C& c, float (&incA)[E], float (&incB)[E], float& uA,
                                          float& uB, uint32_t& sA, uint32_t& sB, unsigned long long pair,
                                          unsigned long long chain_gid, float scale, const float (&dscale)[E]) {
  PairWords<E, PF> pw;
  pw.full(a.rk, c.sub, pair, chain_gid);
  pair_transform<E, IEEE, PF>(a, c, pw.w, incA, incB, uA, uB, sA, sB, scale, dscale);
}

// One uniform per (ladder, sweep, pair) on a separate Philox key.
__device__ __forceinline__ float swap_uniform(unsigned key0, unsigned key1, unsigned long long ladder_gid,
                                              unsigned long long round, int pair) {
  const uint4 r = philox4x32_10((uint32_t)round, ((uint32_t)(round >> 32) & 0xffffu) | ((uint32_t)(pair >> 2) << 16),
                                (uint32_t)ladder_gid, (uint32_t)(ladder_gid >> 32), key0 ^ 0x5851F42Du, key1 ^ 0x4C957F2Du);
  const int q = pair & 3;
  const uint32_t w = q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w));
  return u01_from_bits(w);
}

// Accept rule (rwm_gpu_optimized.py:22-25): (lar > 0) | (u < exp(lar)); NaN compares false -> reject.  In the fast
// path `u` already holds ln(u) (drawn off the critical path), and for u in [0,1) the rule is exactly ln(u) < lar.
template <bool IEEE>
__device__ __forceinline__ bool mh_accept(float lar, float u) {
  if constexpr (IEEE) return (lar > 0.0f) | (u < expf(lar));  // bitwise: no short-circuit branch
  else return u < lar;
}

// pt_rwm_gpu_optimized.py:42-47, literal order; p = min(1, exp(.)), accept iff u < p (:617-621)
template <bool IEEE>
__device__ __forceinline__ bool swap_accept(float bj, float bk, float lj, float lk, float u) {
  using M = Mth<IEEE>;
  const float l = M::sub(M::sub(M::add(M::mul(bj, lk), M::mul(bk, lj)), M::mul(bj, lj)), M::mul(bk, lk));
  const float p = fminf(1.0f, M::exp(l));
  return u < p;  // NaN -> false
}

// One unit of work: the chains of CTA index `cta` (whole ladders) advanced by `n_steps` steps from global step
// `step_offset`.  SLICED: the unit is one time slice of a balanced launch -- another CTA (possibly on another SM) ran
// the previous slice, so state is read past L1 and the accumulators are updated with atomics.
//
// LEAN: two-stage loop for workloads with MANY chains per SM (BASELINE config 5: 16 384 chains of d = 100).  The three-stage
// pipeline below keeps the words of pair p+2, the increments of pair p+1 and the increments of pair p live at once
// (~250 registers at E = 13: two warps per scheduler); the lean loop draws and transforms the words of pair p+1 right
// after the steps of pair p, in place, and is compiled under a register cap so that more warps are resident -- thread
// level parallelism instead of instruction level parallelism.  Same words, same increments, same results.
template <template <int, bool> class Target, int E, bool IEEE, int WT, int PF, bool EXACT, bool TEST, bool SLICED, bool STORE, int LEANMODE = 0>
__device__ __forceinline__ void mcmc_unit(const KernelArgs& a, const long long cta, const long long step_offset,
                                          const long long n_steps, const long long rounds_before) {
  using M = Mth<IEEE>;
  // LEANMODE 1: lean loop, words from PhiloxPairGen (4 registers of run-invariant round state per Philox call);
  // LEANMODE 2: lean loop, ten plain rounds per call (no invariant registers: the leanest form)
  constexpr bool LEAN = LEANMODE != 0;
  extern __shared__ __align__(16) float smem[];  // staged rows are flushed as float4 / float2 vectors

  const int W = WT > 0 ? WT : a.W;
  const int K = a.K, d = a.dim;
  const int tid = (int)threadIdx.x;
  auto cta_sync = [&]() { __syncthreads(); };
  const int cl = tid / W;  // chain within CTA
  CtxT<WT, EXACT> c;
  c.P = a.P; c.d = d; c.W = W;
  c.sub = tid % W;
  c.base = c.sub * E;
  c.lane = tid & 31;
  c.leader = c.lane & ~(W - 1);

  const bool in_cta = cl < a.chains_per_cta;
  const long long chain_raw = cta * a.chains_per_cta + cl;
  const bool valid = in_cta && chain_raw < a.n_chains;
  const long long chain = valid ? chain_raw : 0;  // dummy threads shadow chain 0 but never write
  const int temp = cl % K;  // chains_per_cta is a multiple of K, so this is also chain % K
  const long long ladder = chain / K;
  const bool lead = valid && c.sub == 0;
  const unsigned long long chain_gid = (unsigned long long)(a.chain_id_base + chain);
  const unsigned long long ladder_gid = (unsigned long long)(a.chain_id_base / K + ladder);

  Target<E, IEEE> tgt;
  tgt.init(c);

  float x[E], dscale[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = c.base + e;
    x[e] = (i < d) ? (SLICED ? __ldcg(a.state + chain * d + i) : a.state[chain * d + i]) : 0.0f;
    dscale[e] = i < d ? (a.prop_dim_scale ? a.prop_dim_scale[i] : 1.0f) : 0.0f;  // 0 on padding: Laplace increments stay 0 there
  }
  float lp = SLICED ? __ldcg(a.logp + chain) : a.logp[chain];
  const float beta = a.beta[chain];
  const float scale = a.prop_scale ? a.prop_scale[chain] : 1.0f;

  const long long s_first = step_offset + 1;
  // shared memory carve-up for the swap sweep
  float* s_lp = smem;                                  // [chains_per_cta]
  int* s_src = (int*)(smem + a.chains_per_cta);        // [chains_per_cta]
  int* s_ok = s_src + a.chains_per_cta;                // [chains_per_cta]
  float* s_beta = (float*)(s_ok + a.chains_per_cta);   // [chains_per_cta]
  float* s_x = s_beta + a.chains_per_cta;              // [chains_per_cta, d]
  if (K > 1) {
    if (in_cta && c.sub == 0) s_beta[cl] = beta;
    cta_sync();
  }
  // A ladder whose K*W lanes sit inside one warp (BASELINE config 3: 8 temperatures x 4 lanes = the whole warp) sweeps
  // with warp shuffles: no shared memory, no barrier, no branch.
  const bool warp_ladder = K > 1 && K * W <= 32 && 32 % (K * W) == 0;
  const float beta_next = warp_ladder ? __shfl_down_sync(kFull, beta, W) : 0.0f;

  // countdowns (no per-step modulo)
  long long swap_cd = -1;
  if (K > 1) {
    long long nxt = ((s_first + a.swap_every - 1) / a.swap_every) * a.swap_every;
    if (nxt <= a.burn_in) nxt = (a.burn_in / a.swap_every + 1) * a.swap_every;
    swap_cd = nxt - s_first;
  }
  long long store_cd = -1, store_m = 0;
  // STORE = false: instantiation for runs without retained samples (a.samples == nullptr guaranteed by the launcher): no
  // staging code at all, so the accumulators-only kernels do not depend on the store path
  const bool has_samples = STORE && a.samples != nullptr;
  const bool storing = has_samples && (a.store_mode == RWMPT_STORE_ALL || (a.store_mode == RWMPT_STORE_COLD && temp == 0));
  if (has_samples) {
    const long long r = s_first - a.store_start;
    long long nxt = r <= 0 ? a.store_start + a.thin : s_first + ((a.thin - r % a.thin) % a.thin);
    store_cd = nxt - s_first;
    store_m = (nxt - a.store_start) / a.thin - 1;
  }
  const long long store_chain = a.store_mode == RWMPT_STORE_ALL ? chain : ladder;
  // Retained samples are staged in shared memory, S = a.stage_rows rows per chain, and flushed as contiguous
  // S*d-float blocks (layout (chain, row, dim): the S rows of one chain are adjacent in HBM) with vector stores of
  // a.stage_vw floats (float4 when d % 4 == 0, float2 when d is even): one warp instruction writes 512 contiguous bytes.
  const int S = a.stage_rows;
  // a.stage_bufs == 2: two staging buffers per chain; a full block leaves through the bulk-copy engine (one lane issues
  // cp.async.bulk shared -> global for the chain's S*d*4 contiguous bytes) while the chain stages the next block in the other
  // buffer -- no LDS / STG per lane, no scoreboard wait on the flush.  The engine needs 16-byte aligned source, destination
  // and size: the FIRST block of a chain is cut short (`flush_at`) so that every later block starts on a 16-byte boundary
  // of the sample buffer (d = 50: rows are 200 bytes, every second row is aligned); blocks that still miss the alignment
  // (capacity clipping) take the vector path below.
  // RoughCarpet keeps the single-buffer vector flush AND a single combined (store + accumulators-only) kernel: with the
  // bulk-copy code in it, ptxas schedules the accumulators-only loop of BASELINE config 3 -- the headline workload -- 1.7-3 %
  // slower (2.455e10 vs 2.41e10 / 2.37e10, profiles/r2_variant_ab.txt).  -DRWMPT_RC_BULK=1 -DRWMPT_SPLIT_RC=1 for A/B.
#ifdef RWMPT_RC_BULK
  constexpr bool kBulkOk = true;
#else
  constexpr bool kBulkOk = !(std::is_same<Target<E, IEEE>, RoughCarpetT<E, IEEE, true>>::value || std::is_same<Target<E, IEEE>, RoughCarpetT<E, IEEE, false>>::value);
#endif
  // Staging a row costs one STS per coordinate and little else on every family but RoughCarpet (kFastStage): the lane keeps a
  // pointer to its slot of the current row and stores all E values UNMASKED -- padding coordinates carry exact zeros, and on a
  // padded mapping (d = 50 on 7 x 8) they land in the head of the NEXT row, which the same warp rewrites one step later, or,
  // after the block's last row, in the slack the host leaves behind every buffer (a.stage_stride).  The masked form below spent
  // ~45 instructions per stored step on index arithmetic and predicates (BASELINE config 4: 594 instructions per pair of
  // steps against ~500 without the stores).
  constexpr bool kFastStage = kBulkOk;
  const int st_stride = kFastStage ? a.stage_stride : ((S * d + 3) & ~3);
  const bool bulk = STORE && kBulkOk && a.stage_bufs == 2;
  const int nb = bulk ? 2 : 1;
  float* st_base = smem + a.stage_off;                                        // [chains_per_cta][nb][st_stride]
  float* st_lp_base = st_base + (size_t)a.chains_per_cta * nb * st_stride;     // [chains_per_cta][S]
  float* st_x0 = st_base + (size_t)(in_cta ? cl : 0) * nb * st_stride;
  float* st_x = st_x0;
  int st_buf = 0;
  const bool stage_me = storing && valid;
  // kFastStage: a lane that stores nothing (all padding, a hot chain when only cold chains are retained, a thread without a
  // chain) writes into the slack behind its chain's first buffer instead of branching around the stores -- the slack is
  // never read ((S + 1)*d + E + 1 <= stage_stride: S rows of a block, the overshoot row of stage_put, the slack)
  const bool stage_lane = stage_me && c.base < d;
  float* const st_dummy = st_x0 + (S + 1) * d;
  float* st_row = st_x + c.base;                          // this lane's slot of the row being staged
  float* const st_lp0 = st_lp_base + (size_t)(in_cta ? cl : 0) * (S + 1);   // kFastStage: S + 1 slots per chain (see stage_put)
  float* st_lp_row = st_lp0;                              // the chain's log-density slot of that row
  const bool store_each = has_samples && a.thin == 1 && s_first > a.store_start;  // every step of this run is retained
  int nbuf = 0;
  long long m_base = store_m;   // row index of staged row 0
  int flush_at = S;
  if (bulk) {
    if (stage_me) {
      const unsigned long long addr0 = (unsigned long long)(a.samples + (store_chain * a.sample_stride + m_base) * d);
      const unsigned rb = 4u * (unsigned)d;
      for (int f = 0; f < S; ++f)
        if (((addr0 + (unsigned long long)f * rb) & 15ull) == 0ull) { flush_at = f == 0 ? S : f; break; }
    }
    // the flush holds warp barriers, so every chain of a warp flushes at the same steps: the warp follows its first storing
    // chain (chains whose buffer rows have the other alignment phase -- odd sample_stride with d = 50 -- keep the vector path)
    const unsigned who = __ballot_sync(kFull, stage_me);
    flush_at = __shfl_sync(kFull, flush_at, who ? __ffs(who) - 1 : 0);
  }
  auto stage_flush = [&]() {
    // Each chain's W lanes copy their own chain's staged block: consecutive lanes write consecutive vectors, so every
    // 32-byte sector is written whole exactly once; no CTA barrier (a staging region is only touched by its own warp).
    if (bulk) fence_async_smem();
    __syncwarp();
    long long rows = a.sample_rows - m_base;
    if (rows > nbuf) rows = nbuf;
    if (rows > 0 && stage_me) {
      const int n = (int)rows * d;
      const float* src = st_x;
      float* dst = a.samples + (store_chain * a.sample_stride + m_base) * d;
      if (bulk && (((unsigned long long)dst | (unsigned long long)(4 * n)) & 15ull) == 0ull) {
        if (c.sub == 0) {
          bulk_store_s2g(dst, src, 4u * (unsigned)n);
          bulk_commit();
        }
      } else if (a.stage_vw == 4) {
        for (int v = c.sub; v < (n >> 2); v += W) reinterpret_cast<float4*>(dst)[v] = reinterpret_cast<const float4*>(src)[v];
      } else if (a.stage_vw == 2) {
        for (int v = c.sub; v < (n >> 1); v += W) reinterpret_cast<float2*>(dst)[v] = reinterpret_cast<const float2*>(src)[v];
      } else {
        for (int v = c.sub; v < n; v += W) dst[v] = src[v];
      }
      if (a.sample_logp != nullptr)
        for (int r = c.sub; r < (int)rows; r += W) a.sample_logp[store_chain * a.sample_stride + m_base + r] = kFastStage ? st_lp0[r] : st_lp_base[(size_t)cl * S + r];
    }
    if (bulk) {
      // switch buffers; the block that left from the other buffer one flush ago must have been read completely
      st_buf ^= 1;
      st_x = st_x0 + st_buf * st_stride;
      if (stage_me && c.sub == 0) bulk_wait_read<1>();
    }
    __syncwarp();
    m_base += nbuf;
    nbuf = 0;
    flush_at = S;
    st_row = st_x + c.base;
    st_lp_row = st_lp0;
  };
  auto stage_row = [&](const float (&xs)[E], float lpv) {
    if constexpr (kFastStage) {
      float* const pr = stage_lane ? st_row : st_dummy;
#pragma unroll
      for (int e = 0; e < E; ++e) pr[e] = xs[e];
      float* const pl = (stage_lane && c.sub == 0) ? st_lp_row : st_dummy + E;
      *pl = lpv;
      st_row += d;
      ++st_lp_row;
    } else {
      if (stage_me) {
#pragma unroll
        for (int e = 0; e < E; ++e)
          if (c.ok(e)) st_x[nbuf * d + c.base + e] = xs[e];
        if (c.sub == 0) st_lp_base[(size_t)cl * S + nbuf] = lpv;
      }
    }
    ++nbuf;
    if (kFastStage ? nbuf >= flush_at : nbuf == flush_at) stage_flush();
  };
  // kFastStage, fast loop: the two rows of a pair of steps are staged WITHOUT a flush test in between, so that the pair stays one
  // basic block; the test comes once, after the second row.  A block boundary inside the pair (d = 50: blocks start on even
  // rows, pairs on odd ones) overshoots by one row -- the buffers hold S + 1 rows -- and that row, which is the current
  // state, is staged again as the first row of the next block.
  auto stage_put = [&](const float (&xs)[E], float lpv) {
    float* const pr = stage_lane ? st_row : st_dummy;
#pragma unroll
    for (int e = 0; e < E; ++e) pr[e] = xs[e];
    float* const pl = (stage_lane && c.sub == 0) ? st_lp_row : st_dummy + E;
    *pl = lpv;
    st_row += d;
    ++st_lp_row;
    ++nbuf;
  };
  auto stage_pair_end = [&](const float (&xs)[E], float lpv) {
    if (nbuf >= flush_at) {
      const bool over = nbuf > flush_at;   // warp-uniform, like nbuf and flush_at
      nbuf = flush_at;
      stage_flush();
      if (over) stage_put(xs, lpv);
    }
  };
  const long long burn_t = a.burn_in > step_offset ? a.burn_in - step_offset : 0;  // local steps t >= burn_t count

  unsigned long long n_acc = 0, n_swap_acc = 0, last_attempt = 0;
  unsigned n_acc32 = 0;
  long long round_local = 0;
  float jump_f = 0.0f;
  double jump_d = 0.0;

  // ---- adjacent-temperature sweep of the CTA's ladders (pt_rwm_gpu_optimized.py:594-633); returns whether this
  // thread's chain received a new state ------------------------------------------------------------------------
  // `spare`: this lane's unused accept word of the step the sweep follows (pair_transform).
  auto sweep = [&](const uint32_t spare) -> bool {
    bool changed = false;
    if (warp_ladder && a.swap_mode == RWMPT_SWAP_REFERENCE) {
      // reference semantics (pt_rwm_gpu_optimized.py:594-633 with the copy k -> j of :50-59): pair j only rewrites
      // slot j, from the PRE-sweep occupant of slot j+1 -- every pair decides at once on the pre-sweep values.
      // The pair's uniform is the spare accept word of the colder chain's second lane (its own Philox call when a
      // chain has a single lane).
      const unsigned long long round_g = (unsigned long long)(rounds_before + round_local);  // 0-based
      const bool has_next = valid && temp < K - 1;
      float us;
      if constexpr (TEST) {
        const long long su_base = (round_local * a.n_ladders + ladder) * (K - 1);
        if (a.inj_su != nullptr) us = has_next ? a.inj_su[su_base + temp] : 2.0f;
        else us = W >= 2 ? u01_from_bits(__shfl_sync(kFull, spare, c.leader + 1)) : swap_uniform(a.key0, a.key1, ladder_gid, round_g, temp);
      } else {
        us = W >= 2 ? u01_from_bits(__shfl_sync(kFull, spare, c.leader + 1)) : swap_uniform(a.key0, a.key1, ladder_gid, round_g, temp);
      }
      const float lp_n = __shfl_down_sync(kFull, lp, W);
      float xn[E];
#pragma unroll
      for (int e = 0; e < E; ++e) xn[e] = __shfl_down_sync(kFull, x[e], W);
      const bool ok = has_next && swap_accept<IEEE>(beta, beta_next, lp, lp_n, us);
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = ok ? xn[e] : x[e];
      lp = ok ? lp_n : lp;
      if constexpr (TEST) {
        if (lead && a.swap_dec && has_next) a.swap_dec[(round_local * a.n_ladders + ladder) * (K - 1) + temp] = ok ? 1 : 0;
      }
      n_swap_acc += ok ? 1ull : 0ull;
      last_attempt = ok ? round_g * (unsigned long long)(K - 1) + temp + 1 : last_attempt;
      round_local++;
      return ok;
    }
    if (in_cta) {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (c.base + e < d) s_x[cl * d + c.base + e] = x[e];
      if (c.sub == 0) { s_lp[cl] = lp; s_src[cl] = cl; s_ok[cl] = 0; }
    }
    cta_sync();
    const unsigned long long round_g = (unsigned long long)(rounds_before + round_local);  // 0-based
    const long long su_base = (round_local * a.n_ladders + ladder) * (K - 1);
    const bool inj_su = TEST && a.inj_su != nullptr;
    if (a.swap_mode == RWMPT_SWAP_REFERENCE) {
      // decisions of all pairs are independent here: pair j only ever rewrites slot j
      bool ok = false;
      if (valid && temp < K - 1) {
        const float us = inj_su ? a.inj_su[su_base + temp] : swap_uniform(a.key0, a.key1, ladder_gid, round_g, temp);
        ok = swap_accept<IEEE>(beta, s_beta[cl + 1], s_lp[cl], s_lp[cl + 1], us);
        if (ok) {
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (c.base + e < d) x[e] = s_x[(cl + 1) * d + c.base + e];
          lp = s_lp[cl + 1];
          changed = true;
        }
        if (TEST && lead && a.swap_dec) a.swap_dec[su_base + temp] = ok ? 1 : 0;
      }
      if (ok) { n_swap_acc++; last_attempt = round_g * (unsigned long long)(K - 1) + temp + 1; }
    } else {
      // textbook exchange: sequential sweep by the ladder's first thread, then everyone gathers
      if (valid && temp == 0 && c.sub == 0) {
        for (int j = 0; j < K - 1; ++j) {
          const int sa = s_src[cl + j], sb = s_src[cl + j + 1];
          const float us = inj_su ? a.inj_su[su_base + j] : swap_uniform(a.key0, a.key1, ladder_gid, round_g, j);
          const bool ok = swap_accept<IEEE>(s_beta[cl + j], s_beta[cl + j + 1], s_lp[sa], s_lp[sb], us);
          if (ok) { s_src[cl + j] = sb; s_src[cl + j + 1] = sa; }
          s_ok[cl + j] = ok ? 1 : 0;
          if (TEST && a.swap_dec) a.swap_dec[su_base + j] = ok ? 1 : 0;
        }
      }
      cta_sync();
      if (valid) {
        const int src = s_src[cl];
        if (src != cl) {
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (c.base + e < d) x[e] = s_x[src * d + c.base + e];
          lp = s_lp[src];
          changed = true;
        }
        if (temp < K - 1 && s_ok[cl]) { n_swap_acc++; last_attempt = round_g * (unsigned long long)(K - 1) + temp + 1; }
      }
    }
    round_local++;
    cta_sync();  // s_x / s_lp are rewritten at the next sweep
    return changed;
  };

  // ---- one Metropolis step (+ sweep, accumulators, retained sample) with the given increments -----------------
  auto do_step = [&](const float (&inc)[E], const float u, const long long t, const uint32_t spare) {
    // 2. proposal, 3. its log-density
    float prop[E];
#pragma unroll
    for (int e = 0; e < E; ++e) prop[e] = c.ok(e) ? M::add(x[e], inc[e]) : 0.0f;
    const float lpp = tgt.logp(prop, c);
    // 4. accept rule (rwm_gpu_optimized.py:22-25): NaN compares false -> reject
    const float lar = M::mul(beta, M::sub(lpp, lp));
    const bool acc = mh_accept<IEEE>(lar, u);
    // 5. select
    float xo[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      xo[e] = x[e];
      x[e] = acc ? prop[e] : x[e];
    }
    lp = acc ? lpp : lp;
    const bool post = t >= burn_t;
    n_acc32 += (post && acc) ? 1u : 0u;
    if (TEST && a.decisions != nullptr && lead) a.decisions[t * a.n_chains + chain] = acc ? 1 : 0;

    // 6. adjacent-temperature sweep (whole ladder is in this CTA)
    bool swapped_now = false;
    if (K > 1) {
      if (swap_cd == 0) {
        swap_cd = a.swap_every;
        swapped_now = true;
        sweep(spare);
      }
      swap_cd--;
    }

    // 7. squared jump of this step (Metropolis move + swap move), s > burn_in
    if (IEEE || swapped_now) {
      // exactly the reference's chain[t+1] - chain[t] (rwm_gpu_optimized.py:531)
      float j2 = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float dx = M::sub(x[e], xo[e]);
        j2 = fmaf(dx, dx, j2);
      }
      jump_f += post ? j2 : 0.0f;
    } else {
      float j2 = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) j2 = c.ok(e) ? fmaf(inc[e], inc[e], j2) : j2;
      jump_f += (post && acc) ? j2 : 0.0f;
    }
    // the general path is rare (run edges, burn-in boundary, thinned stores): move the partial sums every time
    jump_d += (double)jump_f; jump_f = 0.0f;
    n_acc += n_acc32; n_acc32 = 0;

    // 8. retained samples: layout (chain, row, dim)
    if (has_samples) {
      if (store_each) {
        stage_row(x, lp);
      } else {
        if (store_cd == 0) {
          store_cd = a.thin;
          stage_row(x, lp);
          store_m++;
        }
        store_cd--;
      }
    }
  };

  // ---- plain step: no sweep, no flush, no burn-in test -- a single basic block, so that ptxas can interleave it with the
  // Philox / Box-Muller stream of the next pair of steps.  Acceptances and squared jumps go to the caller's local sums
  // (`cnt`, `jf`: the fast loop never straddles the burn-in boundary, so it decides once whether they count).  `xo` /
  // `jadd` return the state before the step and the squared jump it added, for the sweep that may follow the second
  // step of a pair.
  auto plain_step = [&](const float (&inc)[E], const float u, auto store_tag, float (&xo)[E], float& jadd, float& jf, unsigned& cnt) {
    float prop[E];
    float j2 = 0.0f;
    // (a Laplace increment is exactly 0 on padding coordinates -- see pair_transform -- so x stays 0 there without a mask)
    constexpr bool kPacked = kUseF32x2 && !IEEE && (EXACT || PF == RWMPT_P_LAPLACE) && E >= 2;
    if constexpr (kPacked) {
      f32x2_t j2p = pack2(0.0f, 0.0f);
#pragma unroll
      for (int e = 0; e + 1 < E; e += 2) {
        const f32x2_t i2 = pack2(inc[e], inc[e + 1]);
        unpack2(add2(pack2(x[e], x[e + 1]), i2), prop[e], prop[e + 1]);
        j2p = fma2(i2, i2, j2p);
      }
      float ja, jb;
      unpack2(j2p, ja, jb);
      j2 = ja + jb;
      if constexpr (E & 1) {
        prop[E - 1] = x[E - 1] + inc[E - 1];
        j2 = fmaf(inc[E - 1], inc[E - 1], j2);
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) prop[e] = c.ok(e) ? M::add(x[e], inc[e]) : 0.0f;
    }
    const float lpp = tgt.logp(prop, c);
    const float lar = M::mul(beta, M::sub(lpp, lp));
    const bool acc = mh_accept<IEEE>(lar, u);
    if constexpr (IEEE) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float xn = acc ? prop[e] : x[e];
        const float dx = M::sub(xn, x[e]);
        j2 = fmaf(dx, dx, j2);
        xo[e] = x[e];
        x[e] = xn;
      }
      jadd = j2;
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if constexpr (!kPacked) j2 = c.ok(e) ? fmaf(inc[e], inc[e], j2) : j2;
        xo[e] = x[e];
        x[e] = acc ? prop[e] : x[e];
      }
      jadd = acc ? j2 : 0.0f;
    }
    jf += jadd;
    lp = acc ? lpp : lp;
    cnt += acc ? 1u : 0u;
    if constexpr (decltype(store_tag)::value) {
      if constexpr (kFastStage) stage_put(x, lp);
      else stage_row(x, lp);
    }
  };

  const bool inject = TEST && a.inj_inc != nullptr;
  if (inject) {
    if constexpr (TEST) {
      for (long long t = 0; t < n_steps; ++t) {
        float inc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = c.base + e;
          inc[e] = (i < d) ? a.inj_inc[(t * a.n_chains + chain) * d + i] : 0.0f;
        }
        do_step(inc, a.inj_u[t * a.n_chains + chain], t, 0u);
      }
    }
  } else if (n_steps > 0) {
    // Software pipeline: the increments of the NEXT pair of steps are drawn while the current pair's density /
    // reduction / accept chain is in flight (they do not depend on the state).  Steps with an "event" (sweep due,
    // sample to retain, accumulator flush, burn-in boundary, start / end of run inside a pair) go through do_step; all
    // other pairs run plain_step twice with no branch at all.
    float iA[E], iB[E], uA, uB;
    uint32_t sA, sB;  // spare accept words of the current pair (swap uniforms)
    unsigned long long pair = (unsigned long long)(s_first - 1) >> 1;
    draw_pair<E, IEEE, PF>(a, c, iA, iB, uA, uB, sA, sB, pair, chain_gid, scale, dscale);
    PhiloxPairGen<PairWords<E, PF>::NC> gen;  // run-invariant part of the Philox rounds
    gen.init(a.rk, c.sub, pair + 1, chain_gid);
    long long t = 0;
    int h0 = (int)((s_first - 1) & 1);  // 1: the run starts on the second step of a pair
    // Sweeps that fall after the second step of a pair (always, when swap_every is even) and the accumulator flush are
    // handled inside the fast loop; only the run's edges, the burn-in boundary, thinned stores and sweeps of an odd
    // swap_every go through the general path.
    const int swap_every = a.swap_every;
    const bool sweep_fast = K > 1 && (swap_every & 1) == 0;
    while (t < n_steps) {
      // first local step >= t that needs the general path
      long long ev = n_steps - 1;                              // last step (also covers an odd tail)
      if (t < burn_t && burn_t - 1 < ev) ev = burn_t - 1;      // burn-in boundary inside a pair
      if (K > 1 && !sweep_fast && t + swap_cd < ev) ev = t + swap_cd;  // sweep due after that step
      if (has_samples && !store_each && t + store_cd < ev) ev = t + store_cd;
      const bool post = t >= burn_t;
      const long long n_fast = h0 ? 0 : (ev - t) >> 1;         // whole pairs strictly before the event
      long long n_done = 0;                                    // pairs the fast loop ran (at most 2^30 per entry)
      auto fast_pairs = [&](auto store_tag) {
        const long long left = n_fast;
        if (left <= 0) return;
        float jf = 0.0f;     // local sums of the squared jumps / acceptances (all steps here are on one side of burn-in)
        unsigned cnt = 0;
        // Three-stage software pipeline: while the steps of pair p run, the Philox words of pair p+1 (drawn one iteration
        // earlier) are turned into increments and the words of pair p+2 are drawn -- three instruction streams with no
        // dependencies between them inside one iteration.
        constexpr int NCW = PairWords<E, PF>::NC;
        constexpr unsigned DD = LEAN ? 1u : 2u;  // how many pairs ahead of the pair being stepped the words are drawn
        uint32_t wn[4 * NCW];
        auto ensure_gen = [&](unsigned long long p) {
          if ((uint32_t)(p >> 32) != gen.hi32) gen.init(a.rk, c.sub, p, chain_gid);
        };
        if constexpr (!LEAN) {
          ensure_gen(pair + 1);
          gen.gen(a.rk, (uint32_t)(pair + 1), wn);
        }
        // The loop runs in chunks that end where something other than a plain pair is due: a sweep after the chunk's last
        // step, the accumulator flush (every chunk end), a change of the high word of the Philox pair counter, or the end
        // of the fast run.
        // (all 32-bit: `left`, `swap_cd` and the Philox low word; the outer loop re-enters for longer runs)
        int left32 = (int)(left < (1ll << 30) ? left : (1ll << 30));
        const int n_here = left32;
        // pairs up to and including the pair the next sweep follows (tracked here only if it can fall inside this entry)
        const bool sw_tracked = sweep_fast && ((swap_cd + 1) >> 1) <= (1ll << 30);
        int sw = sw_tracked ? (int)((swap_cd + 1) >> 1) : 0x7fffffff;
        uint32_t plo = (uint32_t)pair, plo_end;  // low word of the pair being stepped / of the chunk's end
        auto chunk_len = [&]() -> int {
          int m = left32 < RWMPT_CHUNK ? left32 : RWMPT_CHUNK;
          m = m < sw ? m : sw;
          const uint32_t room = 0u - (plo + DD);   // draws before the low word of the Philox counter wraps (0: a full 2^32)
          if (room != 0u && (uint32_t)m > room) m = (int)room;
          return m;
        };
        ensure_gen(pair + DD);
        int len = chunk_len();
        plo_end = plo + (uint32_t)len;
        for (;;) {
          float nA[E], nB[E], vA, vB, xo[E], jadd;
          uint32_t tA, tB;
          if constexpr (LEAN) {
            plain_step(iA, uA, store_tag, xo, jadd, jf, cnt);
            plain_step(iB, uB, std::false_type{}, xo, jadd, jf, cnt);
            if constexpr (LEANMODE == 2) {
              PairWords<E, PF> pw;
              pw.full(a.rk, c.sub, pair + (unsigned long long)(uint32_t)(plo - (uint32_t)pair) + 1ull, chain_gid);
              pair_transform<E, IEEE, PF>(a, c, pw.w, nA, nB, vA, vB, tA, tB, scale, dscale);
            } else {
              gen.gen(a.rk, plo + 1u, wn);
              pair_transform<E, IEEE, PF>(a, c, wn, nA, nB, vA, vB, tA, tB, scale, dscale);
            }
          } else {
#if RWMPT_ORDER == 1
          pair_transform<E, IEEE, PF>(a, c, wn, nA, nB, vA, vB, tA, tB, scale, dscale);
          gen.gen(a.rk, plo + 2u, wn);
          plain_step(iA, uA, store_tag, xo, jadd, jf, cnt);
          plain_step(iB, uB, std::false_type{}, xo, jadd, jf, cnt);
#elif RWMPT_ORDER == 2
          pair_transform<E, IEEE, PF>(a, c, wn, nA, nB, vA, vB, tA, tB, scale, dscale);
          plain_step(iA, uA, store_tag, xo, jadd, jf, cnt);
          gen.gen(a.rk, plo + 2u, wn);
          plain_step(iB, uB, std::false_type{}, xo, jadd, jf, cnt);
#elif RWMPT_ORDER == 3
          plain_step(iA, uA, store_tag, xo, jadd, jf, cnt);
          plain_step(iB, uB, std::false_type{}, xo, jadd, jf, cnt);
          pair_transform<E, IEEE, PF>(a, c, wn, nA, nB, vA, vB, tA, tB, scale, dscale);
          gen.gen(a.rk, plo + 2u, wn);
#else
          plain_step(iA, uA, store_tag, xo, jadd, jf, cnt);
          pair_transform<E, IEEE, PF>(a, c, wn, nA, nB, vA, vB, tA, tB, scale, dscale);
          gen.gen(a.rk, plo + 2u, wn);
          plain_step(iB, uB, std::false_type{}, xo, jadd, jf, cnt);
#endif
          }
          ++plo;
#pragma unroll
          for (int e = 0; e < E; ++e) { iA[e] = nA[e]; iB[e] = nB[e]; }
          uA = vA; uB = vB;
          const uint32_t spareB = sB;
          sA = tA; sB = tB;
          bool done = false;
          if (plo == plo_end) {
            pair += (unsigned)len; left32 -= len; sw -= len;
            if (sw == 0) {  // the sweep is due after the chunk's last step
              sw = swap_every >> 1;
              const bool moved = sweep(spareB);
              // the step's jump is chain[t+1] - chain[t] with the swap included (pt_rwm_gpu_optimized.py:772-789)
              float j2 = 0.0f;
#pragma unroll
              for (int e = 0; e < E; ++e) {
                const float dx = M::sub(x[e], xo[e]);
                j2 = fmaf(dx, dx, j2);
              }
              jf += moved ? j2 - jadd : 0.0f;
            }
            if (post) { jump_d += (double)jf; n_acc += cnt; }  // fp32 partial sums -> fp64 / 64-bit accumulators
            jf = 0.0f; cnt = 0;
            done = left32 == 0;
            if (!done) {
              if (plo + DD < DD) ensure_gen(pair + DD);  // the low word of the next draw wrapped
              len = chunk_len();
              plo_end = plo + (uint32_t)len;
            }
          }
          if constexpr (decltype(store_tag)::value) {  // retained after the sweep, like the reference
            if constexpr (kFastStage) {
              stage_put(x, lp);
              stage_pair_end(x, lp);
            } else {
              stage_row(x, lp);
            }
          }
          if (done) break;
        }
        n_done = n_here;
        if (sweep_fast) swap_cd = sw_tracked ? 2ll * sw - 1 : swap_cd - 2ll * n_here;
      };
      // the retained-sample variant is a separate instantiation so that the accumulators-only loop stays lean
      if constexpr (STORE) {
        if (store_each) fast_pairs(std::true_type{});
        else fast_pairs(std::false_type{});
      } else {
        fast_pairs(std::false_type{});
      }
      t += 2 * n_done;
      if (K > 1 && !sweep_fast) swap_cd -= 2 * n_done;
      if (has_samples) store_cd -= 2 * n_done;
      if (n_done < n_fast) continue;                           // more than 2^30 pairs: re-enter the fast loop
      // the pair that holds the event, through the general path (one copy of do_step: the two halves share the code)
      {
        float nA[E], nB[E], vA, vB;
        uint32_t tA, tB;
        draw_pair<E, IEEE, PF>(a, c, nA, nB, vA, vB, tA, tB, pair + 1, chain_gid, scale, dscale);
#pragma unroll 1
        for (int h = h0; h < 2; ++h) {
          if (t >= n_steps) break;
          float inc[E];
#pragma unroll
          for (int e = 0; e < E; ++e) inc[e] = h ? iB[e] : iA[e];
          do_step(inc, h ? uB : uA, t, h ? sB : sA);
          ++t;
        }
        h0 = 0;
        ++pair;
#pragma unroll
        for (int e = 0; e < E; ++e) { iA[e] = nA[e]; iB[e] = nB[e]; }
        uA = vA; uB = vB;
        sA = tA; sB = tB;
      }
    }
  }

  // epilogue: staged samples, state, log-density, accumulators
  if (has_samples && nbuf > 0) stage_flush();
  if (bulk && stage_me && c.sub == 0) bulk_wait_all<0>();   // every block has left shared memory AND landed before the CTA retires
  jump_d += (double)jump_f;
  n_acc += n_acc32;
  jump_d = group_sum_f64_w<WT>(jump_d, W);
  if (valid) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (c.base + e < d) a.state[chain * d + c.base + e] = x[e];
    if (c.sub == 0) {
      a.logp[chain] = lp;
      if constexpr (SLICED) {
        if (a.accept_count) atomicAdd(a.accept_count + chain, n_acc);
        if (a.sq_jump_sum) atomicAdd(a.sq_jump_sum + chain, jump_d);
        if (K > 1 && temp < K - 1 && a.swap_accepts) atomicAdd(a.swap_accepts + ladder * (K - 1) + temp, n_swap_acc);
        if (K > 1 && a.swap_last_attempt) atomicMax(a.swap_last_attempt + chain, last_attempt);
      } else {
        if (a.accept_count) a.accept_count[chain] += n_acc;
        if (a.sq_jump_sum) a.sq_jump_sum[chain] += jump_d;
        if (K > 1 && temp < K - 1 && a.swap_accepts) a.swap_accepts[ladder * (K - 1) + temp] += n_swap_acc;
        if (K > 1 && a.swap_last_attempt && last_attempt > a.swap_last_attempt[chain]) a.swap_last_attempt[chain] = last_attempt;
      }
    }
  }
}

// Template parameters: Target functor; E coordinates per lane; IEEE parity arithmetic; WT lanes per chain (0 = runtime);
// PF proposal family (-1 = runtime); EXACT: E*W == dim (no padding masks); TEST: injected randomness / decision
// outputs available.
//
// Plain launch (a.n_slices <= 1): CTA b runs unit b for the whole run.  Balanced launch: the run is cut into
// a.n_slices time slices and the grid is sized to the SMs (every scheduler gets the same number of resident warps,
// e.g. 8 one-warp CTAs per SM for the 1024 ladders of BASELINE config 3, where a plain launch leaves 160 of the 592
// schedulers with one warp while the other 432 share two).  CTAs draw tickets g = 0, 1, ... in time-major order
// (slice g / n_units of unit g % n_units), wait -- asleep -- until the unit's previous slice is published, run it, and
// publish.  A ladder is always being advanced by exactly one CTA, the spare CTAs sleep, and the time a warp spends
// alone on its scheduler (where it runs ~1.7x faster) is shared by all ladders instead of ending in an idle tail.
// Results do not depend on the schedule: a slice resumes exactly like a host-level resume (step_offset).
//
// VARIANT (tuned instantiations only): 0 = three-stage pipeline, any CTA size; 1 / 2 = lean loop for one-warp CTAs compiled
// for >= 16 / >= 12 resident CTAs per SM (<= 128 / <= 168 registers: 4 / 3 warps per scheduler); 3 / 4 = the same with ten
// plain Philox rounds per call instead of the PhiloxPairGen invariants (fewer live registers); 5 / 6 = lean loop for CTAs of
// up to 64 threads (stored trajectories allowed), no register cap / <= 168 registers.
__host__ __device__ constexpr int variant_threads(int v) { return v == 0 ? kMaxCtaThreads : (v >= 5 ? 64 : 32); }
__host__ __device__ constexpr int variant_min_ctas(int v) {
  return (v == 1 || v == 3) ? 16 : ((v == 2 || v == 4) ? 12 : (v == 6 ? 6 : 1));
}
__host__ __device__ constexpr int variant_leanmode(int v) { return v == 0 ? 0 : ((v <= 2 || v >= 5) ? 1 : 2); }

template <template <int, bool> class Target, int E, bool IEEE, int WT, int PF, bool EXACT, bool TEST, bool STORE, int VARIANT = 0>
__global__ void __launch_bounds__(variant_threads(VARIANT), variant_min_ctas(VARIANT)) mcmc_kernel(const KernelArgs a) {
  constexpr int LEAN = variant_leanmode(VARIANT);
  if constexpr (TEST) {
    mcmc_unit<Target, E, IEEE, WT, PF, EXACT, TEST, false, STORE>(a, (long long)blockIdx.x, a.step_offset, a.n_steps, a.rounds_before);
  } else {
    // one call site for both schedules (the unit is large and force-inlined): a plain launch is a single "ticket"
    __shared__ unsigned s_ticket;
    const bool sliced = a.n_slices > 1;
    const unsigned total = sliced ? (unsigned)a.n_units * (unsigned)a.n_slices : 0u;
    for (;;) {
      long long unit = (long long)blockIdx.x, so = a.step_offset, n = a.n_steps, rb = a.rounds_before;
      int slice = 0;
      if (sliced) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const unsigned g = s_ticket;
        __syncthreads();
        if (g >= total) break;
        slice = (int)(g / (unsigned)a.n_units);
        unit = (long long)(g % (unsigned)a.n_units);
        if (threadIdx.x == 0) {
          // previous slice published?  Back off up to ~4 us between polls: a slice lasts milliseconds, and a tight poll
          // (the r1j capture shows 2.5 LDG / NANOSLEEP / BRA per ladder-step from waiting CTAs) competes for issue slots
          // with the warp that shares the scheduler
          unsigned ns = 512;
          while (*(volatile int*)(a.unit_done + unit) < slice) {
            __nanosleep(ns);
            ns = ns < 4096 ? ns * 2 : ns;
          }
          __threadfence();
        }
        __syncthreads();
        const long long off = (long long)slice * a.slice_steps;
        n = a.n_steps - off < a.slice_steps ? a.n_steps - off : a.slice_steps;
        so = a.step_offset + off;
        // sweeps performed before global step `so` (host: count_rounds(0, so, burn_in, swap_every))
        rb = (a.K > 1 && so > a.burn_in) ? so / a.swap_every - a.burn_in / a.swap_every : 0;
      }
      // (L1-bypassing loads and atomic accumulators of the sliced unit are harmless in a plain launch)
      mcmc_unit<Target, E, IEEE, WT, PF, EXACT, TEST, true, STORE, LEAN>(a, unit, so, n, rb);
      if (!sliced) break;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicExch(a.unit_done + unit, slice + 1);
    }
  }
}

// ---- batched log-density kernel (rwmpt_log_density) -------------------------------------------
template <template <int, bool> class Target, int E, bool IEEE>
__global__ void __launch_bounds__(kMaxCtaThreads) logp_kernel(const float* __restrict__ P, int d, int W, const float* __restrict__ x,
                                                              long long n, float* __restrict__ out) {
  const int per_cta = blockDim.x / W;
  const int cl = threadIdx.x / W;
  CtxT<0, false> c;
  c.P = P; c.d = d; c.W = W;
  c.sub = threadIdx.x % W;
  c.base = c.sub * E;
  c.lane = threadIdx.x & 31;
  c.leader = c.lane & ~(W - 1);
  Target<E, IEEE> tgt;
  tgt.init(c);
  // two rows per lane group and iteration: both rows' loads are issued before either density is evaluated (the kernel is
  // bound by HBM latency x bytes in flight, not by arithmetic)
  const long long stride = (long long)gridDim.x * per_cta;
  const long long n_round = ((n + 2 * stride - 1) / (2 * stride)) * (2 * stride);  // keep warps converged for the shuffles
  for (long long r = (long long)blockIdx.x * per_cta + cl; r < n_round; r += 2 * stride) {
    const long long r1 = r + stride;
    const bool ok0 = r < n, ok1 = r1 < n;
    float v0[E], v1[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      v0[e] = (ok0 && i < d) ? x[r * d + i] : 0.0f;
      v1[e] = (ok1 && i < d) ? x[r1 * d + i] : 0.0f;
    }
    const float lp0 = tgt.logp(v0, c);
    const float lp1 = tgt.logp(v1, c);
    if (ok0 && c.sub == 0) out[r] = lp0;
    if (ok1 && c.sub == 0) out[r1] = lp1;
  }
}

// Host-side launchers instantiated per target family in rwmpt_inst_*.cu
struct LaunchGeom {
  int E;
  int W;
  int threads;
  int chains_per_cta;
  long long grid;
  size_t smem;
  int sms;       // SM count of the current device
  int schedule;  // RWMPT_SCHEDULE_*
  int variant;   // tuned kernels: loop variant (see mcmc_kernel); 0 unless the geometry table or RWMPT_VARIANT says otherwise
};

}  // namespace rwmpt
