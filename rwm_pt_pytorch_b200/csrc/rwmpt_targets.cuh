// rwmpt_targets.cuh -- target log-densities as device functors, one per family of
// target_distributions/*_torch.py.  A chain's d coordinates are spread over W lanes, E per lane
// (blocked: lane `sub` holds coordinates [sub*E, sub*E+E)); each functor computes its lane-partial
// and reduces over the chain's lanes with XOR butterflies, returning the same value on every lane.
//
// IEEE = true reproduces the reference's fp32 operation order (SURVEY.md section 8a') with un-fused,
// correctly-rounded operations; IEEE = false is the throughput path (MUFU ex2/lg2, FMA, fewer logs).
#pragma once

#include "rwmpt_common.cuh"

namespace rwmpt {

#define RWMPT_NEG_INF (__int_as_float(0xff800000))

// ---- RoughCarpetDistributionTorch.log_density, multimodal_torch.py:470-510 ---------------------
// SCALED = false: instantiated for the tuned kernels when the target carries no per-coordinate scaling factors (the
// common case), so the fast path does not multiply by 1.
template <int E, bool IEEE, bool SCALED>
struct RoughCarpetT {
  using M = Mth<IEEE>;
  float m0, m1, m2, lw0, lw1, lw2, lsp, J;
  float b0, b1, b2, c0, c1, c2;  // fast path: b_k = log2(e) m_k, c_k = log2(e) (lw_k - lsp - m_k^2 / 2)
  bool has_s;
  float s[E];

  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    const float* P = c.P;
    m0 = P[0]; m1 = P[1]; m2 = P[2];
    lw0 = P[3]; lw1 = P[4]; lw2 = P[5];
    lsp = P[6];
    has_s = P[7] != 0.0f;
    J = has_s ? P[8] : 0.0f;
    b0 = kLog2e * m0; b1 = kLog2e * m1; b2 = kLog2e * m2;
    c0 = kLog2e * (lw0 - lsp - 0.5f * m0 * m0); c1 = kLog2e * (lw1 - lsp - 0.5f * m1 * m1);
    c2 = kLog2e * (lw2 - lsp - 0.5f * m2 * m2);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      s[e] = (has_s && i < c.d) ? P[RWMPT_PARAM_HEADER + i] : 1.0f;
    }
  }

#ifndef RWMPT_NO_F32X2
  // Lane partial of the fast path for a mapping without padding (base-2 form, see logp): what one lane contributes before
  // the butterfly over the chain's lanes.  Separate so that a kernel which holds several lanes' coordinates in ONE thread
  // (rwmpt_spec.cuh, one thread per chain) evaluates exactly the same partials and adds them in the butterfly's order.
  __device__ __forceinline__ float lane_part_packed(const float (&x)[E]) const {
    // packed fp32: coordinates (e, e+1) share every FFMA2 / FADD2 / FMUL2; max and ex2 stay scalar
    const f32x2_t B0 = pack2(b0, b0), B1 = pack2(b1, b1), B2 = pack2(b2, b2);
    const f32x2_t C0 = pack2(c0, c0), C1 = pack2(c1, c1), C2 = pack2(c2, c2);
    f32x2_t hs2 = pack2(0.0f, 0.0f), qs2 = pack2(0.0f, 0.0f), ps2 = pack2(1.0f, 1.0f);
#pragma unroll
    for (int e = 0; e + 1 < E; e += 2) {
      f32x2_t xs2 = pack2(x[e], x[e + 1]);
      if constexpr (SCALED) xs2 = mul2(xs2, pack2(s[e], s[e + 1]));
      const f32x2_t l0 = fma2(B0, xs2, C0), l1 = fma2(B1, xs2, C1), l2 = fma2(B2, xs2, C2);
      float l0a, l0b, l1a, l1b, l2a, l2b;
      unpack2(l0, l0a, l0b); unpack2(l1, l1a, l1b); unpack2(l2, l2a, l2b);
      const f32x2_t mx2 = pack2(fmaxf(fmaxf(l0a, l1a), l2a), fmaxf(fmaxf(l0b, l1b), l2b));
      float d0a, d0b, d1a, d1b, d2a, d2b;
      const f32x2_t d0 = sub2(l0, mx2), d1 = sub2(l1, mx2), d2 = sub2(l2, mx2);
      unpack2(d0, d0a, d0b); unpack2(d1, d1a, d1b); unpack2(d2, d2a, d2b);
#ifndef RWMPT_RC_EXP3
      // the largest term is exactly 2^0: two ex2 (smallest and middle exponent) instead of three
      const f32x2_t lo2 = pack2(fminf(fminf(d0a, d1a), d2a), fminf(fminf(d0b, d1b), d2b));
      float mda, mdb, loa, lob;
      unpack2(sub2(add2(add2(d0, d1), d2), lo2), mda, mdb);
      unpack2(lo2, loa, lob);
      const f32x2_t ss2 = add2(add2(pack2(1.0f, 1.0f), pack2(ex2_approx(mda), ex2_approx(mdb))),
                               pack2(ex2_approx(loa), ex2_approx(lob)));
#else
      const f32x2_t ss2 = add2(add2(pack2(ex2_approx(d0a), ex2_approx(d0b)), pack2(ex2_approx(d1a), ex2_approx(d1b))),
                               pack2(ex2_approx(d2a), ex2_approx(d2b)));
#endif
      qs2 = fma2(xs2, xs2, qs2);
      hs2 = add2(hs2, mx2);
      ps2 = mul2(ps2, ss2);
    }
    float hsa, hsb, qsa, qsb, psa, psb;
    unpack2(hs2, hsa, hsb); unpack2(qs2, qsa, qsb); unpack2(ps2, psa, psb);
    if constexpr (E & 1) {
      const float xs = SCALED ? x[E - 1] * s[E - 1] : x[E - 1];
      const float l0 = fmaf(b0, xs, c0), l1 = fmaf(b1, xs, c1), l2 = fmaf(b2, xs, c2);
      const float mx = fmaxf(fmaxf(l0, l1), l2);
#ifndef RWMPT_RC_EXP3
      const float lo = fminf(fminf(l0, l1), l2);
      const float mid = ((l0 + l1) + l2) - (mx + lo);
      const float ss = (1.0f + ex2_approx(mid - mx)) + ex2_approx(lo - mx);
#else
      const float ss = (ex2_approx(l0 - mx) + ex2_approx(l1 - mx)) + ex2_approx(l2 - mx);
#endif
      qsa = fmaf(xs, xs, qsa);
      hsa += mx;
      psa *= ss;
    }
    const float hi = fmaf(-0.5f * kLog2e, qsa + qsb, hsa + hsb);
    const float part = (hi + lg2_approx(psa * psb)) * kLn2;
    return part;
  }
#endif

  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    if constexpr (IEEE) {
      float part = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float xs = has_s ? M::mul(x[e], s[e]) : x[e];
        const float t0 = M::add(M::sub(M::mul(-0.5f, M::sq(M::sub(xs, m0))), lsp), lw0);
        const float t1 = M::add(M::sub(M::mul(-0.5f, M::sq(M::sub(xs, m1))), lsp), lw1);
        const float t2 = M::add(M::sub(M::mul(-0.5f, M::sq(M::sub(xs, m2))), lsp), lw2);
        float mx = fmaxf(fmaxf(t0, t1), t2);
        if (isinf(mx)) mx = 0.0f;  // torch.logsumexp masks infinite maxima
        const float ss = M::add(M::add(M::exp(M::sub(t0, mx)), M::exp(M::sub(t1, mx))), M::exp(M::sub(t2, mx)));
        const float L = M::add(M::log(ss), mx);
        if (c.ok(e)) part = M::add(part, L);
      }
      return M::add(group_sum(part, c), J);
    } else {
      // work in base 2.  t_k = h xs^2 + b_k xs + c_k (h = -log2(e)/2, b_k = log2(e) m_k, c_k = h m_k^2 + a_k): the
      // quadratic term is common to the three modes, so logsumexp_k t_k = h xs^2 + logsumexp_k (b_k xs + c_k);
      // and sum_i log2(S_i) = log2(prod_i S_i): one lg2 per lane instead of one per coordinate.
      // Sums / products are combined as trees (two accumulators) to shorten the dependent chain.  The largest term is exactly
      // 2^0, so two ex2 (smallest and middle exponent) instead of three; -DRWMPT_RC_EXP3 restores the three-ex2 form.
#ifndef RWMPT_NO_F32X2
      if constexpr (C::EXACT && E >= 2) {
        return group_sum(lane_part_packed(x), c) + J;
      }
#endif
      float hs[2] = {0.0f, 0.0f}, ps[2] = {1.0f, 1.0f}, qs[2] = {0.0f, 0.0f};
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float xs = SCALED ? x[e] * s[e] : x[e];
        const float l0 = fmaf(b0, xs, c0), l1 = fmaf(b1, xs, c1), l2 = fmaf(b2, xs, c2);
        const float mx = fmaxf(fmaxf(l0, l1), l2);
#ifndef RWMPT_RC_EXP3
        const float lo = fminf(fminf(l0, l1), l2);
        const float mid = ((l0 + l1) + l2) - (mx + lo);
        const float ss = (1.0f + ex2_approx(mid - mx)) + ex2_approx(lo - mx);
#else
        const float ss = (ex2_approx(l0 - mx) + ex2_approx(l1 - mx)) + ex2_approx(l2 - mx);
#endif
        qs[e & 1] = fmaf(xs, xs, qs[e & 1]);  // padding coordinates hold xs = 0
        hs[e & 1] += c.ok(e) ? mx : 0.0f;
        ps[e & 1] *= c.ok(e) ? ss : 1.0f;
      }
      const float hi = fmaf(-0.5f * kLog2e, qs[0] + qs[1], hs[0] + hs[1]);
      const float prod = ps[0] * ps[1];
      const float part = (hi + lg2_approx(prod)) * kLn2;
      return group_sum(part, c) + J;
    }
  }
};

template <int E, bool IEEE>
using RoughCarpet = RoughCarpetT<E, IEEE, true>;
template <int E, bool IEEE>
using RoughCarpetPlain = RoughCarpetT<E, IEEE, false>;

// ---- ThreeMixtureDistributionTorch.log_density, multimodal_torch.py:173-242 --------------------
template <int E, bool IEEE>
struct ThreeMixture {
  using M = Mth<IEEE>;
  float lw[3], c1[3], J;
  float cst2[3];  // fast path: log2(e) * (c1_k [+ J] + lw_k)
  bool scaled;
  float mu[3][E];  // fast path holds -mu (the centring is one FFMA: x * s - mu)
  float s[E];

  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    const float* P = c.P;
#pragma unroll
    for (int k = 0; k < 3; ++k) { lw[k] = P[k]; c1[k] = P[3 + k]; }
    scaled = P[6] != 0.0f;
    J = P[7];
#pragma unroll
    for (int k = 0; k < 3; ++k) cst2[k] = kLog2e * ((scaled ? c1[k] + J : c1[k]) + lw[k]);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      const bool ok = i < c.d;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float m = ok ? P[RWMPT_PARAM_HEADER + k * c.d + i] : 0.0f;
        mu[k][e] = IEEE ? m : -m;
      }
      s[e] = (ok && scaled) ? P[RWMPT_PARAM_HEADER + 3 * c.d + i] : 1.0f;
    }
  }

  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    if constexpr (!IEEE) {
      // Padding coordinates hold x = 0 and mu = 0 (callers zero them), so no masks.  Base-2 log-sum-exp of the three
      // t_k = -q_k / 2 + const_k; adjacent coordinates share packed FFMA2s.
      float q[3];
#ifndef RWMPT_NO_F32X2
      if constexpr (E >= 2) {
        f32x2_t Q[3] = {pack2(0.0f, 0.0f), pack2(0.0f, 0.0f), pack2(0.0f, 0.0f)};
#pragma unroll
        for (int e = 0; e + 1 < E; e += 2) {
          const f32x2_t X = pack2(x[e], x[e + 1]), S = pack2(s[e], s[e + 1]);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const f32x2_t D = fma2(X, S, pack2(mu[k][e], mu[k][e + 1]));
            Q[k] = fma2(D, D, Q[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float qa, qb;
          unpack2(Q[k], qa, qb);
          q[k] = qa + qb;
          if constexpr (E & 1) {
            const float dd = fmaf(x[E - 1], s[E - 1], mu[k][E - 1]);
            q[k] = fmaf(dd, dd, q[k]);
          }
        }
      } else
#endif
      {
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float dd = fmaf(x[e], s[e], mu[k][e]);
            q[k] = fmaf(dd, dd, q[k]);
          }
        }
      }
      float t[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = fmaf(-0.5f * kLog2e, group_sum(q[k], c), cst2[k]);
      const float mx = fmaxf(fmaxf(fmaxf(t[0], t[1]), t[2]), -3.0e38f);  // all -inf (overflowed state) -> -inf, not NaN
      const float ss = (ex2_approx(t[0] - mx) + ex2_approx(t[1] - mx)) + ex2_approx(t[2] - mx);
      return (lg2_approx(ss) + mx) * kLn2;
    } else {
      float q[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float xs = scaled ? M::mul(x[e], s[e]) : x[e];
        if (c.ok(e)) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float dd = M::sub(xs, mu[k][e]);
            q[k] = M::add(q[k], M::mul(dd, dd));
          }
        }
      }
      float t[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float qq = group_sum(q[k], c);
        float v = M::add(M::mul(-0.5f, qq), c1[k]);
        if (scaled) v = M::add(v, J);
        t[k] = M::add(v, lw[k]);
      }
      float mx = fmaxf(fmaxf(t[0], t[1]), t[2]);
      if (isinf(mx)) mx = 0.0f;
      const float ss = M::add(M::add(M::exp(M::sub(t[0], mx)), M::exp(M::sub(t[1], mx))), M::exp(M::sub(t[2], mx)));
      return M::add(M::log(ss), mx);
    }
  }
};

// ---- FullRosenbrockTorch.log_density, rosenbrock_torch.py:67-84 --------------------------------
template <int E, bool IEEE>
struct FullRosenbrock {
  using M = Mth<IEEE>;
  float a, b;
  float mu[E];

  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    a = c.P[0]; b = c.P[1];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      mu[e] = (i < c.d - 1) ? c.P[RWMPT_PARAM_HEADER + i] : 0.0f;
    }
  }

  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    const float xn_lane = from_next_lane(x[0]);  // x[(sub+1)*E]; masked below when it does not exist
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float xn = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : xn_lane;
      const float r = M::sub(xn, M::sq(x[e]));
      const float t1 = M::mul(b, M::sq(r));
      const float t2 = M::mul(a, M::sq(M::sub(x[e], mu[e])));
      if (c.base + e < c.d - 1) {
        s1 = M::add(s1, t1);
        s2 = M::add(s2, t2);
      }
    }
    if constexpr (IEEE) {
      return -M::add(group_sum(s1, c), group_sum(s2, c));
    } else {
      return -group_sum(s1 + s2, c);
    }
  }
};

// ---- EvenRosenbrockTorch.log_density, rosenbrock_torch.py:194-210 ------------------------------
template <int E, bool IEEE>
struct EvenRosenbrock {
  using M = Mth<IEEE>;
  float a, b;
  float mu[E];

  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    a = c.P[0]; b = c.P[1];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      mu[e] = (i < c.d && (i & 1) == 0) ? c.P[RWMPT_PARAM_HEADER + (i >> 1)] : 0.0f;
    }
  }

  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    const float xn_lane = from_next_lane(x[0]);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      const float xn = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : xn_lane;
      const float t1 = M::mul(a, M::sq(M::sub(x[e], mu[e])));
      const float t2 = M::mul(b, M::sq(M::sub(xn, M::sq(x[e]))));
      if ((i & 1) == 0 && i + 1 < c.d) {
        s1 = M::add(s1, t1);
        s2 = M::add(s2, t2);
      }
    }
    if constexpr (IEEE) {
      return -M::add(group_sum(s1, c), group_sum(s2, c));
    } else {
      return -group_sum(s1 + s2, c);
    }
  }
};

// ---- HybridRosenbrockTorch.log_density, rosenbrock_torch.py:312-351 ----------------------------
template <int E, bool IEEE>
struct HybridRosenbrock {
  using M = Mth<IEEE>;
  float a, b, mu;
  unsigned first_mask;  // bit e set: coordinate is the first of its block (depends on x_0^2)
  unsigned valid_mask;  // bit e set: coordinate index in [1, d)

  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    a = c.P[0]; b = c.P[1]; mu = c.P[2];
    const int n1 = (int)c.P[3];
    const int blk = n1 - 1;
    first_mask = 0; valid_mask = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      if (i >= 1 && i < c.d) {
        valid_mask |= 1u << e;
        if ((i - 1) % blk == 0) first_mask |= 1u << e;
      }
    }
  }

  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    const float x0 = from_leader(x[0], c);
    const float xp_lane = from_prev_lane(x[E - 1]);
    const float x0sq = M::sq(x0);
    float part = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float prev = (e > 0) ? x[(e > 0) ? e - 1 : 0] : xp_lane;
      const float ref = ((first_mask >> e) & 1u) ? x0sq : M::sq(prev);
      const float t = M::mul(b, M::sq(M::sub(x[e], ref)));
      if ((valid_mask >> e) & 1u) part = M::add(part, t);
    }
    const float head = M::mul(-a, M::sq(M::sub(x0, mu)));
    return M::sub(head, group_sum(part, c));
  }
};

// ---- NealFunnelTorch.log_density, funnel_torch.py:39-76 ----------------------------------------
template <int E, bool IEEE>
struct NealFunnel {
  using M = Mth<IEEE>;
  float mu_v, sv, mu_z, lsv, l2p, dm1;

  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    mu_v = c.P[0]; sv = c.P[1]; mu_z = c.P[2]; lsv = c.P[3]; l2p = c.P[4]; dm1 = c.P[5];
  }

  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    const float v = from_leader(x[0], c);
    float q = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      const float dz = M::sub(x[e], mu_z);
      if (i >= 1 && i < c.d) q = IEEE ? M::add(q, M::mul(dz, dz)) : fmaf(dz, dz, q);
    }
    q = group_sum(q, c);
    const float prior = M::sub(M::sub(M::mul(-0.5f, l2p), M::mul(0.5f, lsv)),
                               M::div(M::mul(0.5f, M::sq(M::sub(v, mu_v))), sv));
    if (c.d == 1) return prior;
    const float lik = M::sub(M::sub(M::mul(M::mul(-0.5f, dm1), l2p), M::mul(M::mul(0.5f, dm1), v)),
                             M::mul(M::mul(0.5f, M::exp(-v)), q));
    return M::add(prior, lik);
  }
};

// ---- HypercubeTorch.log_density, hypercube_torch.py:49-80 --------------------------------------
template <int E, bool IEEE>
struct Hypercube {
  float L, R, lud;
  template <class C>
  __device__ __forceinline__ void init(const C& c) { L = c.P[0]; R = c.P[1]; lud = c.P[2]; }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float outside = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (c.ok(e) && !(x[e] >= L && x[e] <= R)) outside += 1.0f;
    return group_sum(outside, c) == 0.0f ? lud : RWMPT_NEG_INF;
  }
};

// ---- IIDGammaTorch.log_density, iid_product_torch.py:52-91 -------------------------------------
template <int E, bool IEEE>
struct IIDGamma {
  using M = Mth<IEEE>;
  float k, th, lnc;
  template <class C>
  __device__ __forceinline__ void init(const C& c) { k = c.P[0]; th = c.P[1]; lnc = c.P[2]; }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float part = 0.0f;
    const float km1 = M::sub(k, 1.0f);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (c.ok(e)) {
        const bool ok = x[e] > 0.0f;
        const float xx = ok ? x[e] : 1.0f;
        const float t = M::sub(M::mul(km1, M::log(xx)), M::div(xx, th));
        part = ok ? M::add(part, t) : RWMPT_NEG_INF;
      }
    }
    return M::sub(group_sum(part, c), lnc);
  }
};

// ---- IIDBetaTorch.log_density, iid_product_torch.py:188-229 ------------------------------------
template <int E, bool IEEE>
struct IIDBeta {
  using M = Mth<IEEE>;
  float al, be, lnc;
  template <class C>
  __device__ __forceinline__ void init(const C& c) { al = c.P[0]; be = c.P[1]; lnc = c.P[2]; }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float part = 0.0f;
    const float am1 = M::sub(al, 1.0f), bm1 = M::sub(be, 1.0f);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (c.ok(e)) {
        const bool ok = x[e] > 0.0f && x[e] < 1.0f;
        const float xx = ok ? x[e] : 0.5f;
        const float t = M::add(M::mul(am1, M::log(xx)), M::mul(bm1, M::log(M::sub(1.0f, xx))));
        part = ok ? M::add(part, t) : RWMPT_NEG_INF;
      }
    }
    return M::add(group_sum(part, c), lnc);
  }
};

// ---- ScaledMultivariateNormalTorch.log_density, multivariate_normal_torch.py:198-223 -----------
template <int E, bool IEEE>
struct ScaledMVN {
  using M = Mth<IEEE>;
  float lnc;
  float cc[E];
  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    lnc = c.P[0];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      cc[e] = (i < c.d) ? c.P[RWMPT_PARAM_HEADER + i] : 0.0f;
    }
  }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float part = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float sx = M::mul(cc[e], x[e]);
      if (c.ok(e)) part = IEEE ? M::add(part, M::mul(sx, sx)) : fmaf(sx, sx, part);
    }
    return M::sub(lnc, M::mul(0.5f, group_sum(part, c)));
  }
};

// ---- MultivariateNormalTorch.log_density with diagonal covariance, multivariate_normal_torch.py:62-92
template <int E, bool IEEE>
struct MVNDiag {
  using M = Mth<IEEE>;
  float lnc;
  float mean[E], prec[E];
  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    lnc = c.P[0];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      mean[e] = (i < c.d) ? c.P[RWMPT_PARAM_HEADER + i] : 0.0f;
      prec[e] = (i < c.d) ? c.P[RWMPT_PARAM_HEADER + c.d + i] : 0.0f;
    }
  }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float part = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float cen = M::sub(x[e], mean[e]);
      const float t = M::mul(M::mul(cen, prec[e]), cen);
      if (c.ok(e)) part = M::add(part, t);
    }
    return M::add(M::mul(-0.5f, group_sum(part, c)), lnc);
  }
};

// ------------------------------------------------------------------------------------------------
// Targets whose density couples every coordinate with every other.  A lane needs the chain's whole
// state: the lanes of the chain all-gather it with shuffles into a per-thread array (dynamic indexing: local memory,
// L1-resident).  Dimension limit kMaxGather.  These are functional rather than tuned: the matvec / data loop per step is
// the cost the reference pays too.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxGather = 128;

template <int E, class C>
__device__ __forceinline__ void gather_all(const float (&v)[E], const C& c, float* all) {
  for (int s = 0; s < c.W; ++s) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float t = __shfl_sync(kFull, v[e], c.leader + s);
      if (s * E + e < kMaxGather) all[s * E + e] = t;
    }
  }
}

// ---- MultivariateNormalTorch.log_density with a general (dense) covariance, multivariate_normal_torch.py:62-92:
// temp = centered @ cov_inv; q = sum(temp * centered); out = -0.5 q + log_norm_const.
// P[0] log norm const; P[16..16+d) mean; then cov_inv row-major (d x d).
template <int E, bool IEEE>
struct MVNDense {
  using M = Mth<IEEE>;
  float lnc;
  float mean[E];
  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    lnc = c.P[0];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      mean[e] = (i < c.d) ? c.P[RWMPT_PARAM_HEADER + i] : 0.0f;
    }
  }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float cen[E], all[kMaxGather];
#pragma unroll
    for (int e = 0; e < E; ++e) cen[e] = c.ok(e) ? M::sub(x[e], mean[e]) : 0.0f;
    gather_all(cen, c, all);
    const float* A = c.P + RWMPT_PARAM_HEADER + c.d;
    float part = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = c.base + e;
      if (i < c.d) {
        float t = 0.0f;
        for (int s = 0; s < c.W; ++s)
          for (int q = 0; q < E; ++q) {
            const int j = s * E + q;                       // coordinate held by lane s, slot q
            if (j < c.d) t = IEEE ? M::add(t, M::mul(all[j], __ldg(A + (size_t)j * c.d + i))) : fmaf(all[j], __ldg(A + (size_t)j * c.d + i), t);
          }
        part = IEEE ? M::add(part, M::mul(t, cen[e])) : fmaf(t, cen[e], part);
      }
    }
    return M::add(M::mul(-0.5f, group_sum(part, c)), lnc);
  }
};

// ---- SuperFunnelTorch.log_density, funnel_torch.py:193-291 (hierarchical logistic regression).
// theta = (alpha[J], beta[J][K], mu_alpha, mu_beta[K], tau_alpha, tau_beta).
// P[0] J, P[1] K, P[2] prior hyper-mean variance, P[3] its log, P[4] prior tau scale, P[5] its log, P[6] log 2pi, P[7] log 2,
// P[8] log pi, P[9] number of observations N; then N records of (group j, y, x[K]) after the header.
template <int E, bool IEEE>
struct SuperFunnel {
  using M = Mth<IEEE>;
  int J, K, N;
  float hv, lhv, ts, lts, l2p, l2, lpi;
  template <class C>
  __device__ __forceinline__ void init(const C& c) {
    const float* P = c.P;
    J = (int)P[0]; K = (int)P[1]; hv = P[2]; lhv = P[3]; ts = P[4]; lts = P[5]; l2p = P[6]; l2 = P[7]; lpi = P[8]; N = (int)P[9];
  }
  static __device__ __forceinline__ float log_sigmoid(float z) {   // F.logsigmoid: min(z, 0) - log1p(exp(-|z|))
    const float a = fabsf(z);
    return IEEE ? fminf(z, 0.0f) - log1pf(expf(-a)) : fminf(z, 0.0f) - M::log(1.0f + M::exp(-a));
  }
  template <class C>
  __device__ __forceinline__ float logp(const float (&x)[E], const C& c) const {
    float th[kMaxGather];
    gather_all(x, c, th);
    const int o_ma = J + J * K, o_mb = o_ma + 1, o_ta = o_mb + K, o_tb = o_ta + 1;
    const float mu_a = th[o_ma], tau_a = th[o_ta], tau_b = th[o_tb];
    // likelihood: the chain's lanes share the observations (:236-249)
    const float* D = c.P + RWMPT_PARAM_HEADER;
    float ll = 0.0f;
    for (int n = c.sub; n < N; n += c.W) {
      const float* rec = D + (size_t)n * (K + 2);
      const int j = (int)rec[0];
      const float y = rec[1];
      float eta = th[j];
      for (int k = 0; k < K; ++k) eta = fmaf(rec[2 + k], th[J + j * K + k], eta);
      ll += y * log_sigmoid(eta) + (1.0f - y) * log_sigmoid(-eta);
    }
    ll = group_sum(ll, c);
    if (!(tau_a > 1e-9f) || !(tau_b > 1e-9f)) return RWMPT_NEG_INF;                       // :225-227
    // priors, evaluated by every lane on the gathered state (identical values on all lanes) (:256-291)
    float pa = 0.0f, pb = 0.0f;
    const float ita2 = 1.0f / (tau_a * tau_a), itb2 = 1.0f / (tau_b * tau_b), lta = M::log(tau_a), ltb = M::log(tau_b);
    for (int j = 0; j < J; ++j) {
      const float da = th[j] - mu_a;
      pa += -0.5f * l2p - lta - 0.5f * da * da * ita2;
      float sq = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float db = th[J + j * K + k] - th[o_mb + k];
        sq = fmaf(db, db, sq);
      }
      pb += -0.5f * (float)K * l2p - (float)K * ltb - 0.5f * sq * itb2;
    }
    float smb = 0.0f;
    for (int k = 0; k < K; ++k) smb = fmaf(th[o_mb + k], th[o_mb + k], smb);
    const float p_ma = -0.5f * l2p - 0.5f * lhv - 0.5f * mu_a * mu_a / hv;
    const float p_mb = -0.5f * (float)K * l2p - 0.5f * (float)K * lhv - 0.5f * smb / hv;
    const float ra = tau_a / ts, rb = tau_b / ts;
    const float p_ta = l2 - lpi - lts - (IEEE ? log1pf(ra * ra) : M::log(1.0f + ra * ra));
    const float p_tb = l2 - lpi - lts - (IEEE ? log1pf(rb * rb) : M::log(1.0f + rb * rb));
    return ((((((ll + pa) + pb) + p_ma) + p_mb) + p_ta) + p_tb);
  }
};

}  // namespace rwmpt
