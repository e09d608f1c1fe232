// rwmpt_spec.cuh -- warp-specialised RWM / PT-RWM kernel for the FEW-WARPS regime: BASELINE config 3 under strong scaling
// (1024 ladders sharded over 8 GPUs leave 128 ladders = 128 warps on a GPU with 592 warp schedulers) and BASELINE config 2
// (4096 chains of d = 20 are 512 warps).
//
// There the fused kernel is bound by the latency of ONE warp: a step is ~100-200 dependent-ish instructions, most of them
// randomness (Philox rounds, Box-Muller) that does not depend on the chain at all, while schedulers idle.  Here a warp's worth
// of chains (a whole ladder for PT) is a CTA of 1 + NP warps on as many schedulers:
//   * the NP PRODUCER warps draw the Philox words of chunks of step pairs and turn them into scaled increments, log-uniforms
//     and spare swap words -- exactly the words and transforms of mcmc_kernel (PhiloxPairGen, pair_transform), lane for lane --
//     and leave them in a shared-memory ring (one float4-interleaved slot per lane and pair, two buffers per producer);
//   * the CONSUMER warp reads its lanes' slots and does nothing but propose / evaluate / accept / sweep.
// Hand-over is by named barriers (bar.sync / bar.arrive on a full / empty pair per ring buffer): no CTA-wide barrier, no
// polling.  A chunk ends where a sweep is due, so the consumer sweeps (warp shuffles, as in mcmc_kernel) while the producers
// are already filling other buffers.
//
// Results are those of mcmc_kernel: same Philox counters, same arithmetic, so states, log-densities, acceptance and swap
// counters are bit-identical and the squared-jump sums agree to the grouping of their fp32 partial sums
// (tests/test_gpu_parity.py::test_specialised_*).  The kernel takes only the regular part of a run -- an even number of steps
// from an even offset, all on one side of the burn-in boundary, swap_every even, chains that fill whole warps without padding
// (for PT: a ladder that fills exactly one warp; config 2's d = 30 runs on 8 x 4 lanes with two masked padding coordinates),
// accumulators only; the host (rwmpt_api.cu) runs the edges through
// mcmc_kernel, which is resumable by construction.
#pragma once

#include "rwmpt_kernel.cuh"

namespace rwmpt {

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// geometry of the ring: SLOT float4 per lane and pair (2E increments, 2 log-uniforms, 2 spare words), NB = 2 NP buffers of CH
// pairs each, at most 48 KiB in all
template <int E, int NP>
struct SpecRing {
  static constexpr int SLOT = (2 * E + 4 + 3) / 4;
  static constexpr int NB = 2 * NP;
  static constexpr int CH_RAW = (48 * 1024) / (NB * SLOT * 512);   // static shared memory: 48 KiB
  static constexpr int CH = CH_RAW < 1 ? 1 : (CH_RAW > 8 ? 8 : CH_RAW);
};

// Shared chunk arithmetic of all roles: pairs in the next chunk given the pairs left and the pairs up to and including the one
// the next sweep follows (`sw`, 0x7fffffff when the launch has no sweeps).
template <int CH>
__device__ __forceinline__ int spec_chunk_len(long long left, int sw) {
  const int m = left < CH ? (int)left : CH;
  return m < sw ? m : sw;
}

// CW = consumer lanes per chain: WT (the consumer keeps the fused kernel's lane mapping) or 1 (ONE consumer thread per chain: it
// holds all E*WT coordinates, evaluates the WT lane partials of the density itself -- independent instruction streams instead
// of a shuffle butterfly -- and adds them in the butterfly's order, so the log-density is bit-identical; measured 2x slower
// (issue-bound), kept as a knob for the RoughCarpet 5 x 4 shape).  NP = producer warps.  EXACT = false: E * WT > dim, the last
// lanes of a chain carry padding coordinates (config 2's d = 30 on 8 x 4) -- masked exactly as in mcmc_unit.
template <template <int, bool> class Target, int E, int WT, int PF, int CW, int NP, bool EXACT = true>
__global__ void __launch_bounds__(32 * (1 + NP)) mcmc_spec_kernel(const KernelArgs a) {
  using R = SpecRing<E, NP>;
  constexpr int SLOT = R::SLOT, NB = R::NB, CH = R::CH;
  constexpr int CPW = 32 / WT;                                // chains per warp
  static_assert(CW == WT || CW == 1, "consumer lanes per chain: WT or 1");
  static_assert(CW == WT || WT == 4, "the one-thread-per-chain consumer is written for four producer lanes per chain");
  static_assert(CW == WT || EXACT, "the one-thread-per-chain consumer has no padding masks");
  constexpr bool IEEE = false;
  using M = Mth<IEEE>;
  __shared__ float4 ring[NB][CH][SLOT][32];                   // [buffer][pair][word group][slot]: conflict-free LDS.128 / STS.128

  const int K = a.K, d = a.dim;
  const int lane = (int)threadIdx.x & 31;
  const int role = (int)threadIdx.x >> 5;                   // 0: consumer (steps), 1 .. NP: producers (randomness)
  CtxT<WT, EXACT> c;
  c.P = a.P; c.d = d; c.W = WT;
  c.sub = lane % WT;
  c.base = c.sub * E;
  c.lane = lane;
  c.leader = lane & ~(WT - 1);
  const int cw = lane / WT;                                   // chain within the warp
  const long long chain = (long long)blockIdx.x * CPW + cw;   // K > 1: the warp is one ladder (K == CPW), cw is the temperature
  const long long ladder = chain / K;
  const int temp = K > 1 ? cw : 0;
  const unsigned long long chain_gid = (unsigned long long)(a.chain_id_base + chain);

  const unsigned long long pair0 = (unsigned long long)a.step_offset >> 1;      // step_offset is even
  const long long n_pairs = a.n_steps >> 1;                                      // n_steps is even
  const bool post = a.step_offset >= a.burn_in;                                  // the whole launch is on one side of burn-in
  const bool sweeps = post && K > 1;
  const int half_se = a.swap_every >> 1;                                         // swap_every is even
  // pairs up to and including the pair the first sweep follows: sweeps follow global steps s = m * swap_every > burn_in
  int sw0 = 0x7fffffff;
  if (sweeps) {
    const long long s_first = a.step_offset + 1;
    long long nxt = ((s_first + a.swap_every - 1) / a.swap_every) * a.swap_every;
    if (nxt <= a.burn_in) nxt = (a.burn_in / a.swap_every + 1) * a.swap_every;
    const long long pairs_to = (nxt - a.step_offset) >> 1;                       // nxt and step_offset are even
    sw0 = pairs_to < 0x7fffffff ? (int)pairs_to : 0x7fffffff;
  }

  if (role >= 1) {
    // ------------------------------------------------ producer ------------------------------------------------
    // Producer pi fills chunks pi, pi + NP, pi + 2 NP, ... into its own two buffers pi and pi + NP (alternating), each with its
    // own full / empty barrier pair shared with the consumer only.
    const int pi = role - 1;
    const float scale = a.prop_scale ? a.prop_scale[chain] : 1.0f;
    float dscale[E];
#pragma unroll
    for (int e = 0; e < E; ++e) dscale[e] = c.ok(e) ? (a.prop_dim_scale ? a.prop_dim_scale[c.base + e] : 1.0f) : 0.0f;
    PhiloxPairGen<PairWords<E, PF>::NC> gen;
    unsigned long long pair = pair0;
    gen.init(a.rk, c.sub, pair, chain_gid);
    long long left = n_pairs;
    int sw = sw0;
    long long chunk_no = 0, mine = 0;
    while (left > 0) {
      const int len = spec_chunk_len<CH>(left, sw);
      if ((int)(chunk_no % NP) == pi) {
        const int buf = pi + NP * (int)(mine & 1);
        if (mine >= 2) named_bar_sync(NB + buf, 64);                              // buffer `buf` has been consumed
        for (int p = 0; p < len; ++p) {
          const unsigned long long pr = pair + (unsigned)p;
          if ((uint32_t)(pr >> 32) != gen.hi32) gen.init(a.rk, c.sub, pr, chain_gid);
          uint32_t w[4 * PairWords<E, PF>::NC];
          gen.gen(a.rk, (uint32_t)pr, w);
          float iA[E], iB[E], uA, uB;
          uint32_t sA, sB;
          pair_transform<E, IEEE, PF>(a, c, w, iA, iB, uA, uB, sA, sB, scale, dscale);
          float v[4 * SLOT];
#pragma unroll
          for (int q = 0; q < 4 * SLOT; ++q) v[q] = 0.0f;
#pragma unroll
          for (int e = 0; e < E; ++e) { v[e] = iA[e]; v[E + e] = iB[e]; }
          v[2 * E] = uA; v[2 * E + 1] = uB;
          v[2 * E + 2] = __uint_as_float(sA); v[2 * E + 3] = __uint_as_float(sB);
          // CW == WT: slot = producer lane (the consumer lane of the same index reads it back); CW == 1: slot = sub * CPW + chain,
          // so that the consumer lanes (one per chain) read consecutive float4
          const int slot = CW == 1 ? c.sub * CPW + cw : lane;
#pragma unroll
          for (int g = 0; g < SLOT; ++g) ring[buf][p][g][slot] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        named_bar_arrive(buf, 64);                                                // buffer `buf` is full
        ++mine;
      }
      pair += (unsigned)len; left -= len;
      if (sweeps) sw -= len;
      if (sw == 0) sw = half_se;
      ++chunk_no;
    }
    // match the consumer's last "empty" arrivals so that no barrier is left half-armed when the CTA retires
    if (mine >= 2) named_bar_sync(NB + pi + NP * (int)(mine & 1), 64);
    if (mine >= 1) named_bar_sync(NB + pi + NP * (int)((mine - 1) & 1), 64);
    return;
  }

  // -------------------------------------------------- consumer --------------------------------------------------
  if constexpr (CW == 1) {
    constexpr int ET = E * WT;                                  // all coordinates of a chain in one thread
    const int kc = lane < K ? lane : K - 1;                     // lanes >= K shadow the hottest chain and never write
    const bool active = lane < K;
    const long long ch = (long long)blockIdx.x * CPW + kc;
    Target<E, IEEE> tgt;                                        // the E-coordinate functor: evaluated WT times per density
    tgt.init(c);
    float x[ET];
#pragma unroll
    for (int i = 0; i < ET; ++i) x[i] = a.state[ch * d + i];
    float lp = a.logp[ch];
    const float beta = a.beta[ch];
    const float beta_next = __shfl_down_sync(kFull, beta, 1);
    unsigned long long n_acc = 0, n_swap_acc = 0, last_attempt = 0;
    double jump_d = 0.0;
    long long round_local = 0;
    auto density = [&](const float (&y)[ET]) -> float {
      float part[WT];
#pragma unroll
      for (int v = 0; v < WT; ++v) {
        float seg[E];
#pragma unroll
        for (int e = 0; e < E; ++e) seg[e] = y[v * E + e];
        part[v] = tgt.lane_part_packed(seg);
      }
      // group_sum_w<4>: xor 2 then xor 1, as seen from lane 0 -- every lane of the butterfly ends with the same bits
      return __fadd_rn(__fadd_rn(part[0], part[2]), __fadd_rn(part[1], part[3])) + tgt.J;
    };
    auto step1 = [&](const float (&inc)[ET], const float u, float (&xo)[ET], float& jadd, float& jf, unsigned& cnt) {
      float prop[ET];
      float j2 = 0.0f;
#pragma unroll
      for (int i = 0; i < ET; ++i) {
        prop[i] = x[i] + inc[i];
        j2 = fmaf(inc[i], inc[i], j2);
      }
      const float lpp = density(prop);
      const float lar = M::mul(beta, M::sub(lpp, lp));
      const bool acc = mh_accept<IEEE>(lar, u);
#pragma unroll
      for (int i = 0; i < ET; ++i) {
        xo[i] = x[i];
        x[i] = acc ? prop[i] : x[i];
      }
      jadd = acc ? j2 : 0.0f;
      jf += jadd;
      lp = acc ? lpp : lp;
      cnt += acc ? 1u : 0u;
    };
    long long left = n_pairs, chunk_no = 0;
    int sw = sw0;
    while (left > 0) {
      const int len = spec_chunk_len<CH>(left, sw);
      const int buf = (int)(chunk_no % NB);
      named_bar_sync(buf, 64);
      float jf = 0.0f;
      unsigned cnt = 0;
      float xo[ET], jadd = 0.0f;
      uint32_t spareB = 0u;
      for (int p = 0; p < len; ++p) {
        float iA[ET], iB[ET], uA = 0.0f, uB = 0.0f;
#pragma unroll
        for (int v = 0; v < WT; ++v) {
          float w[4 * SLOT];
#pragma unroll
          for (int g = 0; g < SLOT; ++g) {
            const float4 t = ring[buf][p][g][v * CPW + kc];
            w[4 * g] = t.x; w[4 * g + 1] = t.y; w[4 * g + 2] = t.z; w[4 * g + 3] = t.w;
          }
#pragma unroll
          for (int e = 0; e < E; ++e) { iA[v * E + e] = w[e]; iB[v * E + e] = w[E + e]; }
          if (v == 0) { uA = w[2 * E]; uB = w[2 * E + 1]; }                 // the chain's log-uniforms (leader's words)
          if (v == 1) spareB = __float_as_uint(w[2 * E + 3]);               // spare word of the chain's second lane
        }
        step1(iA, uA, xo, jadd, jf, cnt);
        step1(iB, uB, xo, jadd, jf, cnt);
      }
      named_bar_arrive(NB + buf, 64);
      left -= len;
      ++chunk_no;
      if (sweeps) sw -= len;
      if (sw == 0) {
        sw = half_se;
        const unsigned long long round_g = (unsigned long long)(a.rounds_before + round_local);
        const bool has_next = active && kc < K - 1;
        const float us = u01_from_bits(spareB);
        const float lp_n = __shfl_down_sync(kFull, lp, 1);
        float xn[ET];
#pragma unroll
        for (int i = 0; i < ET; ++i) xn[i] = __shfl_down_sync(kFull, x[i], 1);
        const bool ok = has_next && swap_accept<IEEE>(beta, beta_next, lp, lp_n, us);
        float j2 = 0.0f;
#pragma unroll
        for (int i = 0; i < ET; ++i) {
          x[i] = ok ? xn[i] : x[i];
          const float dx = M::sub(x[i], xo[i]);
          j2 = fmaf(dx, dx, j2);
        }
        lp = ok ? lp_n : lp;
        n_swap_acc += ok ? 1ull : 0ull;
        last_attempt = ok ? round_g * (unsigned long long)(K - 1) + kc + 1 : last_attempt;
        round_local++;
        jf += ok ? j2 - jadd : 0.0f;
      }
      if (post) { jump_d += (double)jf; n_acc += cnt; }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < ET; ++i) a.state[ch * d + i] = x[i];
      a.logp[ch] = lp;
      if (a.accept_count) a.accept_count[ch] += n_acc;
      if (a.sq_jump_sum) a.sq_jump_sum[ch] += jump_d;
      if (K > 1 && kc < K - 1 && a.swap_accepts) a.swap_accepts[ladder * (K - 1) + kc] += n_swap_acc;
      if (K > 1 && a.swap_last_attempt && last_attempt > a.swap_last_attempt[ch]) a.swap_last_attempt[ch] = last_attempt;
    }
    return;
  }
  Target<E, IEEE> tgt;
  tgt.init(c);
  float x[E];
#pragma unroll
  for (int e = 0; e < E; ++e) x[e] = c.ok(e) ? a.state[chain * d + c.base + e] : 0.0f;
  float lp = a.logp[chain];
  const float beta = a.beta[chain];
  const float beta_next = __shfl_down_sync(kFull, beta, WT);
  const bool lead = c.sub == 0;
  unsigned long long n_acc = 0, n_swap_acc = 0, last_attempt = 0;
  double jump_d = 0.0;
  long long round_local = 0;

  // one Metropolis step on the lane's coordinates: the arithmetic of mcmc_unit's plain_step (packed fp32 on an EXACT mapping,
  // scalar with the padding masks otherwise)
  auto step = [&](const float (&inc)[E], const float u, float (&xo)[E], float& jadd, float& jf, unsigned& cnt) {
    float prop[E];
    float j2 = 0.0f;
    if constexpr (EXACT && kUseF32x2 && E >= 2) {
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
      for (int e = 0; e < E; ++e) {
        prop[e] = c.ok(e) ? M::add(x[e], inc[e]) : 0.0f;
        j2 = c.ok(e) ? fmaf(inc[e], inc[e], j2) : j2;
      }
    }
    const float lpp = tgt.logp(prop, c);
    const float lar = M::mul(beta, M::sub(lpp, lp));
    const bool acc = mh_accept<IEEE>(lar, u);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      xo[e] = x[e];
      x[e] = acc ? prop[e] : x[e];
    }
    jadd = acc ? j2 : 0.0f;
    jf += jadd;
    lp = acc ? lpp : lp;
    cnt += acc ? 1u : 0u;
  };

  long long left = n_pairs, chunk_no = 0;
  int sw = sw0;
  while (left > 0) {
    const int len = spec_chunk_len<CH>(left, sw);
    const int buf = (int)(chunk_no % NB);
    named_bar_sync(buf, 64);                                                      // wait until buffer `buf` is full
    float jf = 0.0f;
    unsigned cnt = 0;
    float xo[E], jadd = 0.0f;
    uint32_t spareB = 0u;
    for (int p = 0; p < len; ++p) {
      float v[4 * SLOT];
#pragma unroll
      for (int g = 0; g < SLOT; ++g) {
        const float4 t = ring[buf][p][g][lane];
        v[4 * g] = t.x; v[4 * g + 1] = t.y; v[4 * g + 2] = t.z; v[4 * g + 3] = t.w;
      }
      float iA[E], iB[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { iA[e] = v[e]; iB[e] = v[E + e]; }
      spareB = __float_as_uint(v[2 * E + 3]);
      step(iA, v[2 * E], xo, jadd, jf, cnt);
      step(iB, v[2 * E + 1], xo, jadd, jf, cnt);
    }
    named_bar_arrive(NB + buf, 64);                                               // buffer `buf` may be refilled
    left -= len;
    ++chunk_no;
    if (sweeps) sw -= len;
    if (sw == 0) {
      // adjacent-temperature sweep after the chunk's last step: reference semantics on the pre-sweep values, the pair's
      // uniform is the spare accept word of the colder chain's second lane (mcmc_unit::sweep, warp-ladder form)
      sw = half_se;
      const unsigned long long round_g = (unsigned long long)(a.rounds_before + round_local);
      const bool has_next = temp < K - 1;
      const float us = u01_from_bits(__shfl_sync(kFull, spareB, c.leader + 1));
      const float lp_n = __shfl_down_sync(kFull, lp, WT);
      float xn[E];
#pragma unroll
      for (int e = 0; e < E; ++e) xn[e] = __shfl_down_sync(kFull, x[e], WT);
      const bool ok = has_next && swap_accept<IEEE>(beta, beta_next, lp, lp_n, us);
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = ok ? xn[e] : x[e];
      lp = ok ? lp_n : lp;
      n_swap_acc += ok ? 1ull : 0ull;
      last_attempt = ok ? round_g * (unsigned long long)(K - 1) + temp + 1 : last_attempt;
      round_local++;
      // the step's jump is chain[t+1] - chain[t] with the swap included (pt_rwm_gpu_optimized.py:772-789)
      float j2 = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float dx = M::sub(x[e], xo[e]);
        j2 = fmaf(dx, dx, j2);
      }
      jf += ok ? j2 - jadd : 0.0f;
    }
    if (post) { jump_d += (double)jf; n_acc += cnt; }
  }

  jump_d = group_sum_f64_w<WT>(jump_d, WT);
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (c.ok(e)) a.state[chain * d + c.base + e] = x[e];
  if (lead) {
    a.logp[chain] = lp;
    if (a.accept_count) a.accept_count[chain] += n_acc;
    if (a.sq_jump_sum) a.sq_jump_sum[chain] += jump_d;
    if (K > 1 && temp < K - 1 && a.swap_accepts) a.swap_accepts[ladder * (K - 1) + temp] += n_swap_acc;
    if (K > 1 && a.swap_last_attempt && last_attempt > a.swap_last_attempt[chain]) a.swap_last_attempt[chain] = last_attempt;
  }
}

// grid: one CTA per warp's worth of chains; the caller guarantees n_chains % (32 / WT) == 0 and E * WT == dim (EXACT) or
// E * WT >= dim (otherwise)
template <template <int, bool> class Target, int E, int WT, int PF, int CW, int NP, bool EXACT = true>
cudaError_t launch_mcmc_spec(const KernelArgs& a, cudaStream_t st) {
  const long long grid = a.n_chains / (32 / WT);
  mcmc_spec_kernel<Target, E, WT, PF, CW, NP, EXACT><<<(unsigned)grid, 32 * (1 + NP), 0, st>>>(a);
  return cudaGetLastError();
}

// Per-family entry points, defined in the family's fast translation unit; cudaErrorNotSupported when the shape (elements per
// lane, lanes per chain, consumer lanes, producers; Normal proposal) has no instantiation.
cudaError_t launch_spec_rough_carpet(const KernelArgs& a, int E, int W, int consumer_lanes, int producers, cudaStream_t st);
cudaError_t launch_spec_even_rosenbrock(const KernelArgs& a, int E, int W, int consumer_lanes, int producers, cudaStream_t st);

}  // namespace rwmpt
