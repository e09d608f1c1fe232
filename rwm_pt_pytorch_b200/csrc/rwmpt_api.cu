// rwmpt_api.cu -- the extern "C" boundary of librwmpt.so (see include/rwmpt.h): argument validation,
// launch geometry, family dispatch, and the small stand-alone kernels (proposal sampler, PT swap sweep,
// ESJD reduction, Philox known-answer hook) plus the host-buffer end-to-end entry.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "rwmpt_launch.cuh"
#include "rwmpt_spec.cuh"

namespace rwmpt {

static thread_local char g_err[512] = "";

// warp-specialised kernel for RWM (config 2), from measurements (profiles/r2_specialised_kernel.txt): 4096 chains of d = 10
// (256 fused warps) +57 % with three producer warps per consumer warp, d = 20 (512 fused warps) +7 % with two, d = 30 on 8 x 4
// (512 fused warps, two padding coordinates) +37 % with two (1.41e10 -> 1.93e10; three: the same, four: -2 %); more
// producers lose again (the machine's integer-multiply / issue capacity, not the consumer, is the limit).  Auto rule: up
// to 4 fused-kernel warps per SM.
constexpr int kSpecRwmAutoWarpsPerSm = 4;
static int spec_rwm_producers(int elems_per_lane, int lanes_per_chain) { return elems_per_lane == 8 ? 2 : (lanes_per_chain == 2 ? 3 : 2); }

// family name (RWMPT_FAMILY_LIST) -> enum of include/rwmpt.h
#define RWMPT_FAMILY_ID_rough_carpet RWMPT_T_ROUGH_CARPET
#define RWMPT_FAMILY_ID_three_mixture RWMPT_T_THREE_MIXTURE
#define RWMPT_FAMILY_ID_full_rosenbrock RWMPT_T_FULL_ROSENBROCK
#define RWMPT_FAMILY_ID_even_rosenbrock RWMPT_T_EVEN_ROSENBROCK
#define RWMPT_FAMILY_ID_hybrid_rosenbrock RWMPT_T_HYBRID_ROSENBROCK
#define RWMPT_FAMILY_ID_neal_funnel RWMPT_T_NEAL_FUNNEL
#define RWMPT_FAMILY_ID_hypercube RWMPT_T_HYPERCUBE
#define RWMPT_FAMILY_ID_iid_gamma RWMPT_T_IID_GAMMA
#define RWMPT_FAMILY_ID_iid_beta RWMPT_T_IID_BETA
#define RWMPT_FAMILY_ID_scaled_mvn RWMPT_T_SCALED_MVN
#define RWMPT_FAMILY_ID_mvn_diag RWMPT_T_MVN_DIAG
#define RWMPT_FAMILY_ID_mvn_dense RWMPT_T_MVN_DENSE
#define RWMPT_FAMILY_ID_super_funnel RWMPT_T_SUPER_FUNNEL

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int cuda_fail(cudaError_t e, const char* what) {
  return fail(RWMPT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

static const int kFastE[] = {
#define X(e) e,
    RWMPT_FAST_E_LIST(X)
#undef X
};
static const int kIeeeE[] = {
#define X(e) e,
    RWMPT_IEEE_E_LIST(X)
#undef X
};

static int64_t min_params(int family, int d) {
  const int64_t H = RWMPT_PARAM_HEADER;
  switch (family) {
    case RWMPT_T_ROUGH_CARPET: return H;  // + d when P[7] != 0 (checked by the facade; device reads only then)
    case RWMPT_T_THREE_MIXTURE: return H + 3LL * d;
    case RWMPT_T_FULL_ROSENBROCK: return H + d - 1;
    case RWMPT_T_EVEN_ROSENBROCK: return H + d / 2;
    case RWMPT_T_SCALED_MVN: return H + d;
    case RWMPT_T_MVN_DIAG: return H + 2LL * d;
    case RWMPT_T_MVN_DENSE: return H + d + (int64_t)d * d;
    default: return H;
  }
}

static int check_target(const rwmpt_target_t* t) {
  if (!t) return fail(RWMPT_EINVAL, "target is NULL");
  if (t->family < 0 || t->family >= RWMPT_T_COUNT) return fail(RWMPT_EINVAL, "unknown target family %d", t->family);
  if (t->dim < 1) return fail(RWMPT_EINVAL, "dim must be >= 1 (got %d)", t->dim);
  if (!t->params) return fail(RWMPT_EINVAL, "target params pointer is NULL");
  if (t->n_params < min_params(t->family, t->dim))
    return fail(RWMPT_EINVAL, "target family %d with dim %d needs >= %lld params, got %lld", t->family, t->dim,
                (long long)min_params(t->family, t->dim), (long long)t->n_params);
  if (t->family == RWMPT_T_EVEN_ROSENBROCK && (t->dim < 2 || t->dim % 2))
    return fail(RWMPT_EINVAL, "EvenRosenbrock needs an even dim >= 2");
  if ((t->family == RWMPT_T_FULL_ROSENBROCK) && t->dim < 2) return fail(RWMPT_EINVAL, "FullRosenbrock needs dim >= 2");
  if ((t->family == RWMPT_T_MVN_DENSE || t->family == RWMPT_T_SUPER_FUNNEL) && t->dim > kMaxGather)
    return fail(RWMPT_ENOTSUP, "target family %d gathers the whole state on every lane: dim <= %d", t->family, kMaxGather);
  return RWMPT_OK;
}

// Choose lanes-per-chain W and elements-per-lane E.  Candidates: E from the compiled list, W a power of two,
// E*W >= d, no lane entirely padding, and the ladder (K*W threads) must fit one CTA.  Prefer the least padding;
// among equals prefer more lanes while the grid is too small to fill the machine (latency hiding), else fewer.
static int pick_geometry(int d, int K, long long n_ladders, bool ieee, int want_W, LaunchGeom* g, int family = -1, int pf = -1,
                         bool for_mcmc = false) {
  const int* list = ieee ? kIeeeE : kFastE;
  const int n_list = ieee ? (int)(sizeof(kIeeeE) / sizeof(int)) : (int)(sizeof(kFastE) / sizeof(int));
  const long long n_chains = n_ladders * K;
  const double fill_threads = 148.0 * 4 * 32 * 4;  // ~4 warps per scheduler
  double best_score = 1e300;
  int bestE = -1, bestW = -1;
  // two passes: the register-friendly elements-per-lane counts first (E <= 13); E = 25 only when nothing else fits (a long
  // ladder whose n_temps * lanes would exceed one CTA, or a dimension above 13 * 32)
  for (int pass = 0; pass < 2 && bestE < 0; ++pass)
  for (int W = 1; W <= 32; W *= 2) {
    if (want_W > 0 && W != want_W) continue;
    if ((long long)K * W > kMaxCtaThreads) continue;
    for (int k = 0; k < n_list; ++k) {
      const int E = list[k];
      if (pass == 0 && E > 13) break;
      if ((long long)E * W < d) continue;
      if (want_W <= 0 && W > 1 && (long long)(W - 1) * E >= d) continue;  // auto mode: no lane that is all padding
      const double waste = (double)E * W / d;
      const double threads = (double)n_chains * W;
      double score = waste;
      if (threads < fill_threads) score *= 1.0 + 0.15 * (fill_threads / threads > 8 ? 3.0 : (fill_threads / threads - 1.0) * 3.0 / 7.0);
      score += 1e-3 * E;  // mild preference for fewer registers
      if (score < best_score) { best_score = score; bestE = E; bestW = W; }
      break;  // list is ascending: first E that fits is the least padded for this W
    }
  }
  // tuned-only geometry: ThreeMixture d = 50..56 with a Laplace / UniformRadius proposal (BASELINE config 4) runs 7
  // coordinates x 8 lanes (rwmpt_inst_three_mixture.cu); E = 7 is not in the generic lists
  if (!ieee && want_W <= 0 && family == RWMPT_T_THREE_MIXTURE && (pf == RWMPT_P_LAPLACE || pf == RWMPT_P_UNIFORM_RADIUS) &&
      d > 49 && d <= 56 && K * 8 <= kMaxCtaThreads) {
    bestE = 7;
    bestW = 8;
  }
  // BASELINE config 2 at d = 30 (any 24 < d <= 32): EvenRosenbrock RWM with a Normal proposal runs the tuned 8 x 4 kernel -- the
  // score above would take 4 x 8 for 4096 chains; measured 1.41e10 against 1.14e10 chain-steps/s fused, and 8 x 4 is the shape the
  // warp-specialised kernel is instantiated for (1.93e10, profiles/r2_specialised_kernel.txt)
  if (!ieee && for_mcmc && want_W <= 0 && family == RWMPT_T_EVEN_ROSENBROCK && pf == RWMPT_P_NORMAL && K == 1 && d > 24 && d <= 32) {
    bestE = 8;
    bestW = 4;
  }
  g->variant = 0;
  if (!ieee && for_mcmc) {
    // tuning knobs for A/B measurements (results never depend on the geometry): RWMPT_GEOM="E,W" forces a tuned
    // (elements per lane, lanes per chain) pair that a family's translation unit instantiates; RWMPT_VARIANT the loop variant
    const char* eg = getenv("RWMPT_GEOM");
    int fe = 0, fw = 0;
    if (eg && sscanf(eg, "%d,%d", &fe, &fw) == 2 && fe >= 1 && fw >= 1 && fw <= 32 && (fw & (fw - 1)) == 0 &&
        (long long)fe * fw >= d && (long long)K * fw <= kMaxCtaThreads) {
      bestE = fe;
      bestW = fw;
    }
    // measured defaults (profiles/r2_variant_ab.txt): with thousands of one-warp units per GPU the lean loop under a register
    // cap beats the three-stage pipeline -- BASELINE config 5, d = 100 on 13 x 8: FullRosenbrock +23 % with <= 128
    // registers (4 warps per scheduler), NealFunnel +13 % with <= 168 (3 per scheduler)
    if (bestE == 13 && bestW == 8 && pf == RWMPT_P_NORMAL && K == 1 && n_chains * bestW / 32 >= 148LL * 12) {
      if (family == RWMPT_T_FULL_ROSENBROCK) g->variant = 1;
      if (family == RWMPT_T_NEAL_FUNNEL) g->variant = 2;
    }
    // BASELINE config 4 (ThreeMixture d = 50 on 7 x 8, two warps per ladder, trajectories stored): the lean loop issues ~8 %
    // fewer instructions per step (no copies between pipeline stages) at the same occupancy
    if (family == RWMPT_T_THREE_MIXTURE && bestE == 7 && bestW == 8) g->variant = 5;
    const char* ev = getenv("RWMPT_VARIANT");
    if (ev) g->variant = atoi(ev);
  }
  if (bestE < 0)
    return fail(RWMPT_ENOTSUP, "no kernel variant for dim=%d n_temps=%d lanes_per_chain=%d (max dim %d; n_temps*lanes <= %d)",
                d, K, want_W, list[n_list - 1] * 32, kMaxCtaThreads);
  g->E = bestE;
  g->W = bestW;
  const int ladder_threads = K * bestW;
  int ladders_per_cta = ladder_threads >= 32 ? 1 : 32 / ladder_threads;
  g->chains_per_cta = ladders_per_cta * K;
  g->threads = ((g->chains_per_cta * bestW + 31) / 32) * 32;
  g->grid = (n_ladders + ladders_per_cta - 1) / ladders_per_cta;
  g->smem = K > 1 ? (size_t)g->chains_per_cta * (4 + d) * sizeof(float) : 0;
  return RWMPT_OK;
}

static int64_t count_rounds(int64_t step_offset, int64_t n_steps, int64_t burn_in, int32_t swap_every) {
  // sweeps at global steps s in (step_offset, step_offset + n_steps] with s % swap_every == 0 and s > burn_in
  if (swap_every < 1 || n_steps <= 0) return 0;
  const int64_t lo = step_offset > burn_in ? step_offset : burn_in;  // s > lo
  const int64_t hi = step_offset + n_steps;
  if (hi <= lo) return 0;
  return hi / swap_every - lo / swap_every;
}

static cudaError_t dispatch_mcmc(int family, const KernelArgs& a, const LaunchGeom& g, bool ieee, cudaStream_t st) {
  switch (family) {
    case RWMPT_T_ROUGH_CARPET: return launch_mcmc_rough_carpet(a, g, ieee, st);
    case RWMPT_T_THREE_MIXTURE: return launch_mcmc_three_mixture(a, g, ieee, st);
    case RWMPT_T_FULL_ROSENBROCK: return launch_mcmc_full_rosenbrock(a, g, ieee, st);
    case RWMPT_T_EVEN_ROSENBROCK: return launch_mcmc_even_rosenbrock(a, g, ieee, st);
    case RWMPT_T_HYBRID_ROSENBROCK: return launch_mcmc_hybrid_rosenbrock(a, g, ieee, st);
    case RWMPT_T_NEAL_FUNNEL: return launch_mcmc_neal_funnel(a, g, ieee, st);
    case RWMPT_T_HYPERCUBE: return launch_mcmc_hypercube(a, g, ieee, st);
    case RWMPT_T_IID_GAMMA: return launch_mcmc_iid_gamma(a, g, ieee, st);
    case RWMPT_T_IID_BETA: return launch_mcmc_iid_beta(a, g, ieee, st);
    case RWMPT_T_SCALED_MVN: return launch_mcmc_scaled_mvn(a, g, ieee, st);
    case RWMPT_T_MVN_DIAG: return launch_mcmc_mvn_diag(a, g, ieee, st);
    case RWMPT_T_MVN_DENSE: return launch_mcmc_mvn_dense(a, g, ieee, st);
    case RWMPT_T_SUPER_FUNNEL: return launch_mcmc_super_funnel(a, g, ieee, st);
  }
  return cudaErrorInvalidValue;
}

static cudaError_t dispatch_logp(int family, const float* P, int d, int E, int W, const float* x, long long n, float* out,
                                 bool ieee, cudaStream_t st) {
  switch (family) {
    case RWMPT_T_ROUGH_CARPET: return launch_logp_rough_carpet(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_THREE_MIXTURE: return launch_logp_three_mixture(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_FULL_ROSENBROCK: return launch_logp_full_rosenbrock(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_EVEN_ROSENBROCK: return launch_logp_even_rosenbrock(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_HYBRID_ROSENBROCK: return launch_logp_hybrid_rosenbrock(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_NEAL_FUNNEL: return launch_logp_neal_funnel(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_HYPERCUBE: return launch_logp_hypercube(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_IID_GAMMA: return launch_logp_iid_gamma(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_IID_BETA: return launch_logp_iid_beta(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_SCALED_MVN: return launch_logp_scaled_mvn(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_MVN_DIAG: return launch_logp_mvn_diag(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_MVN_DENSE: return launch_logp_mvn_dense(P, d, E, W, x, n, out, ieee, st);
    case RWMPT_T_SUPER_FUNNEL: return launch_logp_super_funnel(P, d, E, W, x, n, out, ieee, st);
  }
  return cudaErrorInvalidValue;
}

static cudaError_t dispatch_swap_prob(int family, const float* P, int d, int E, int W, float bc, float bs, long long n, unsigned k0,
                                      unsigned k1, long long row_base, double* sum_out, cudaStream_t st) {
  switch (family) {
#define X(name, cls) \
  case RWMPT_FAMILY_ID_##name: return launch_swap_prob_##name(P, d, E, W, bc, bs, n, k0, k1, row_base, sum_out, st);
    RWMPT_FAMILY_LIST(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

static int run_impl(const rwmpt_run_args_t* r, void* stream, bool require_rwm) {
  if (!r) return fail(RWMPT_EINVAL, "args is NULL");
  int rc = check_target(&r->target);
  if (rc) return rc;
  const int d = r->target.dim;
  if (r->n_temps < 1) return fail(RWMPT_EINVAL, "n_temps must be >= 1");
  if (require_rwm && r->n_temps != 1) return fail(RWMPT_EINVAL, "rwmpt_rwm_run needs n_temps == 1 (use rwmpt_pt_run)");
  if (r->n_ladders < 0 || r->n_steps < 0 || r->burn_in < 0 || r->step_offset < 0)
    return fail(RWMPT_EINVAL, "n_ladders, n_steps, burn_in and step_offset must be >= 0");
  if (r->proposal_family < RWMPT_P_NORMAL || r->proposal_family > RWMPT_P_UNIFORM_RADIUS)
    return fail(RWMPT_EINVAL, "unknown proposal family %d", r->proposal_family);
  if (r->math_mode != RWMPT_MATH_FAST && r->math_mode != RWMPT_MATH_IEEE) return fail(RWMPT_EINVAL, "bad math_mode");
  if (r->swap_mode != RWMPT_SWAP_REFERENCE && r->swap_mode != RWMPT_SWAP_EXCHANGE) return fail(RWMPT_EINVAL, "bad swap_mode");
  if (r->n_temps > 1 && r->swap_every < 1) return fail(RWMPT_EINVAL, "swap_every must be >= 1");
  if (!r->state || !r->logp || !r->beta) return fail(RWMPT_EINVAL, "state, logp and beta must be non-NULL");
  const bool inject = r->inj_increments != nullptr;
  if (inject != (r->inj_uniforms != nullptr)) return fail(RWMPT_EINVAL, "inj_increments and inj_uniforms go together");
  if (!inject && !r->prop_scale) return fail(RWMPT_EINVAL, "prop_scale is required unless increments are injected");
  // a run with injected increments has no Philox stream of its own (the facade passes seed 0): the sweep's uniforms must
  // be injected too, otherwise the two sweep forms (shuffle / shared memory) would fall back to different sources
  if (inject && r->n_temps > 1 && !r->inj_swap_uniforms && count_rounds(r->step_offset, r->n_steps, r->burn_in, r->swap_every) > 0)
    return fail(RWMPT_EINVAL, "inj_increments on a ladder (n_temps > 1) needs inj_swap_uniforms as well");
  if (r->samples) {
    if (r->store_mode != RWMPT_STORE_COLD && r->store_mode != RWMPT_STORE_ALL)
      return fail(RWMPT_EINVAL, "samples given but store_mode is NONE");
    if (r->thin < 1 || r->sample_stride < 1 || r->store_start < 0 || r->sample_rows < 0 || r->sample_rows > r->sample_stride)
      return fail(RWMPT_EINVAL, "bad thin / sample_stride / sample_rows / store_start");
  }
  if (r->lanes_per_chain < 0 || r->lanes_per_chain > 32 || (r->lanes_per_chain & (r->lanes_per_chain - 1)))
    return fail(RWMPT_EINVAL, "lanes_per_chain must be 0 or a power of two <= 32");
  const bool test_mode = inject || r->inj_swap_uniforms || r->decisions || r->swap_decisions;
  if (test_mode && r->math_mode != RWMPT_MATH_IEEE)
    return fail(RWMPT_ENOTSUP, "injected randomness / decision outputs (test mode) need math_mode = RWMPT_MATH_IEEE");
  if (r->n_ladders == 0 || r->n_steps == 0) return RWMPT_OK;  // empty input: nothing to do

  LaunchGeom g;
  const bool ieee = r->math_mode == RWMPT_MATH_IEEE;
  rc = pick_geometry(d, r->n_temps, r->n_ladders, ieee, r->lanes_per_chain, &g, r->target.family, r->proposal_family, true);
  if (rc) return rc;
  if (g.grid > 2147483647LL) return fail(RWMPT_ENOTSUP, "too many CTAs (%lld)", g.grid);
  if (r->schedule < RWMPT_SCHEDULE_AUTO || r->schedule > RWMPT_SCHEDULE_SPECIALISED)
    return fail(RWMPT_EINVAL, "unknown schedule %d", r->schedule);
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
    g.sms = sms;
    g.schedule = r->schedule;
    // injected randomness / decision outputs (test mode) always run the plain schedule
    if (r->inj_increments || r->decisions || r->swap_decisions) g.schedule = RWMPT_SCHEDULE_PLAIN;
  }

  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.P = r->target.params; a.dim = d; a.prop_family = r->proposal_family; a.K = r->n_temps; a.W = g.W;
  a.chains_per_cta = g.chains_per_cta;
  a.swap_every = r->n_temps > 1 ? r->swap_every : 1;
  a.swap_mode = r->swap_mode; a.store_mode = r->store_mode;
  a.prop_scale = r->prop_scale; a.prop_dim_scale = r->prop_dim_scale; a.beta = r->beta;
  a.n_ladders = r->n_ladders; a.n_chains = r->n_ladders * r->n_temps;
  a.n_steps = r->n_steps; a.burn_in = r->burn_in; a.step_offset = r->step_offset;
  a.rounds_before = count_rounds(0, r->step_offset, r->burn_in, a.swap_every);
  a.state = r->state; a.logp = r->logp;
  a.key0 = (unsigned)(r->seed & 0xffffffffu); a.key1 = (unsigned)(r->seed >> 32);
  for (int q = 0; q < 10; ++q) {
    a.rk[2 * q] = a.key0 + (unsigned)q * 0x9E3779B9u;
    a.rk[2 * q + 1] = a.key1 + (unsigned)q * 0xBB67AE85u;
  }
  a.chain_id_base = r->chain_id_base;
  a.target_plain = r->target.n_params < RWMPT_PARAM_HEADER + d ? 1 : 0;
  a.samples = r->samples; a.sample_logp = r->sample_logp;
  a.store_start = r->store_start; a.thin = r->thin < 1 ? 1 : r->thin; a.sample_stride = r->sample_stride; a.sample_rows = r->sample_rows;
  a.accept_count = r->accept_count; a.sq_jump_sum = r->sq_jump_sum;
  a.swap_accepts = r->swap_accepts; a.swap_last_attempt = r->swap_last_attempt;
  a.inj_inc = r->inj_increments; a.inj_u = r->inj_uniforms; a.inj_su = r->inj_swap_uniforms;
  a.decisions = r->decisions; a.swap_dec = r->swap_decisions;

  if (r->samples) {
    // staging region after the swap region: S rows per chain (as many as keep the CTA's staging under ~32 KiB, several
    // CTAs per SM), double-buffered for the bulk-copy flush of the fast kernels (RWMPT_BULK_STORE=0 restores the
    // single-buffer vector flush for A/B measurements)
    const size_t swap_floats = (g.smem / sizeof(float) + 3) & ~(size_t)3;
    const char* eb = getenv("RWMPT_BULK_STORE");
    const int bufs = (!ieee && !(eb && atoi(eb) == 0)) ? 2 : 1;
    // every family but RoughCarpet stages rows with unmasked stores (mcmc_unit, kFastStage): up to E - 1 padding zeros spill past
    // a buffer's last row, into a slack of 28 floats (E <= 25)
    // ... and the buffers hold one row more than a block (the fast loop tests for a full block once per pair of steps)
    const bool fast_stage = r->target.family != RWMPT_T_ROUGH_CARPET;
    const size_t slack = fast_stage ? (size_t)d + 28 : 0;
    // Staging budget per CTA: 32 KiB, or up to 64 KiB when the launch has so few CTAs per SM that all of them stay resident with
    // the larger blocks (BASELINE config 4: 512 CTAs on 148 SMs, 16 rows instead of 8 per flush: 7.35e9 -> 7.83e9 chain-steps/s;
    // 4 rows: 6.63e9).  RWMPT_STAGE_KIB overrides for A/B measurements.
    const long long ctas_per_sm = (g.grid + (g.sms > 0 ? g.sms : 148) - 1) / (g.sms > 0 ? g.sms : 148);
    const size_t per_cta = (size_t)227 * 1024 / (size_t)(ctas_per_sm < 1 ? 1 : (ctas_per_sm > 32 ? 32 : ctas_per_sm)) - 1024;
    const size_t fixed = swap_floats * sizeof(float) + (size_t)g.chains_per_cta * 17 * sizeof(float) + 64;
    size_t stage_cap = per_cta > fixed + 32 * 1024 ? per_cta - fixed : 32 * 1024;
    if (stage_cap > 64 * 1024) stage_cap = 64 * 1024;
    const char* ek = getenv("RWMPT_STAGE_KIB");
    if (ek && atoi(ek) >= 4 && atoi(ek) <= 160) stage_cap = (size_t)atoi(ek) * 1024;
    int bufs_eff = bufs;
    auto stage_bytes = [&](int rows) { return (size_t)bufs_eff * g.chains_per_cta * (((size_t)rows * d + slack + 3) & ~(size_t)3) * sizeof(float); };
    int S = 16;
    if (fast_stage) {
      // an even number of rows, at least two: blocks then end where the next one starts 16-byte aligned whenever d is even
      while (S > 2 && stage_bytes(S) > stage_cap) S -= 2;
      if (stage_bytes(S) > 160 * 1024) bufs_eff = 1;   // a very wide CTA: single buffer, vector flush
    } else {
      while (S > 1 && stage_bytes(S) > stage_cap) S >>= 1;
    }
    const size_t st_stride = ((size_t)S * d + slack + 3) & ~(size_t)3;
    const size_t lp_floats = ((size_t)g.chains_per_cta * (S + 1) + 1) & ~(size_t)1;
    a.stage_rows = S;
    a.stage_bufs = bufs_eff;
    a.stage_stride = (int)st_stride;
    a.stage_off = (int)swap_floats;
    const uintptr_t p = reinterpret_cast<uintptr_t>(r->samples);
    a.stage_vw = (d % 4 == 0 && p % 16 == 0) ? 4 : ((d % 2 == 0 && p % 8 == 0) ? 2 : 1);
    g.smem = (swap_floats + (size_t)bufs_eff * g.chains_per_cta * st_stride + lp_floats) * sizeof(float);
  }
  // Few warps per GPU (strong scaling of config 3, config 2's 4096 chains): the warp-specialised kernel (rwmpt_spec.cuh) takes the
  // regular middle of the run, mcmc_kernel the edges -- up to the first even step at or past burn-in, and an odd last step --
  // each launch resuming the previous one exactly (state, log-density, accumulators and Philox offsets are all functions of
  // step_offset).  Shapes: chains fill whole warps (E * W == dim, or EvenRosenbrock 24 < d <= 32 on 8 x 4), Normal proposal, accumulators only; for a
  // ladder: 8 temperatures x 4 lanes = one warp, swap_every even, the reference's swap semantics.
  // Auto: PT up to 3.5 ladders per SM (measured: +40 % at 64-296 ladders per GPU, +8 % at 512, -24 % at 1024); RWM see below.
  // RWMPT_SCHEDULE_SPECIALISED forces it where eligible; RWMPT_SPEC_CW / RWMPT_SPEC_NP choose the consumer mapping / producers.
  const long long n_chains_all = r->n_ladders * r->n_temps;
  const int cpw = 32 / g.W;
  const bool spec_common = !ieee && !test_mode && !r->samples && r->proposal_family == RWMPT_P_NORMAL &&
                           n_chains_all % cpw == 0 && n_chains_all / cpw <= 2147483647LL;
  const bool spec_exact = g.E * g.W == d;
  const bool spec_pt = spec_common && spec_exact && r->target.family == RWMPT_T_ROUGH_CARPET && a.target_plain && r->n_temps == 8 && g.W == 4 && g.E == 5 &&
                       (a.swap_every & 1) == 0 && r->swap_mode == RWMPT_SWAP_REFERENCE;
  const bool spec_rwm = spec_common && r->target.family == RWMPT_T_EVEN_ROSENBROCK && r->n_temps == 1 &&
                        ((spec_exact && g.E == 5 && (g.W == 4 || g.W == 2)) || (g.E == 8 && g.W == 4 && d > 24 && d <= 32) ||
                         (g.E == 4 && g.W == 8 && d > 28 && d <= 32));   // d = 30: two padding coordinates
  const long long fused_warps = n_chains_all / cpw;
  bool spec_want = r->schedule == RWMPT_SCHEDULE_SPECIALISED;
  if (r->schedule == RWMPT_SCHEDULE_AUTO && g.sms > 0) {
    if (spec_pt) spec_want = fused_warps * 2 <= 7LL * g.sms;
    if (spec_rwm) spec_want = fused_warps <= (long long)kSpecRwmAutoWarpsPerSm * g.sms;
  }
  if ((spec_pt || spec_rwm) && spec_want) {
    const char* ecw = getenv("RWMPT_SPEC_CW");
    const char* enp = getenv("RWMPT_SPEC_NP");
    const int spec_cw = ecw ? atoi(ecw) : g.W;
    const int spec_np = enp ? atoi(enp) : (spec_rwm ? spec_rwm_producers(g.E, g.W) : 1);
    const int64_t O = r->step_offset, N = r->n_steps, B = r->burn_in;
    int64_t head = O >= B ? 0 : B - O;          // steps that end at burn-in ...
    if ((O + head) & 1) ++head;                 // ... or one later, so that the middle starts on an even step
    if (head > N) head = N;
    const int64_t mid = (N - head) & ~(int64_t)1;
    if (mid >= 64) {
      const int64_t segs[3][2] = {{O, head}, {O + head, mid}, {O + head + mid, N - head - mid}};
      for (int k = 0; k < 3; ++k) {
        if (segs[k][1] <= 0) continue;
        KernelArgs s = a;
        s.step_offset = segs[k][0];
        s.n_steps = segs[k][1];
        s.rounds_before = count_rounds(0, s.step_offset, r->burn_in, a.swap_every);
        cudaError_t e = cudaErrorNotSupported;
        if (k == 1)
          e = spec_pt ? launch_spec_rough_carpet(s, g.E, g.W, spec_cw, spec_np, (cudaStream_t)stream)
                      : launch_spec_even_rosenbrock(s, g.E, g.W, spec_cw, spec_np, (cudaStream_t)stream);
        // no instantiation for this shape / producer count (e.g. an RWMPT_SPEC_NP outside 1..4): the fused kernel runs the middle too
        if (e == cudaErrorNotSupported) e = dispatch_mcmc(r->target.family, s, g, ieee, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, k == 1 ? "specialised mcmc kernel launch" : "mcmc kernel launch");
      }
      return RWMPT_OK;
    }
  }
  cudaError_t e = dispatch_mcmc(r->target.family, a, g, ieee, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "mcmc kernel launch");
  return RWMPT_OK;
}

// ---- proposal sampler (plugin sample(n)): a group of G lanes per row, G = smallest power of two >= ceil(d / 4) (at most
// a warp), each lane turning one Philox call into four coordinates and writing them as one float4 when the row layout
// allows.  Counter = (block of four coordinates, tag, row id): counter-based, so the ball sampler's second pass
// regenerates the normals instead of staging them.  Same fast transforms as the fused kernel.
__global__ void __launch_bounds__(128) proposal_kernel(int family, int d, int G, float scale, const float* __restrict__ dscale,
                                                       long long n, unsigned k0, unsigned k1, long long row_base, int vec4,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const int rows_per_warp = 32 / G;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int n_blk = (d + 3) / 4;
  const long long n_iter = (n + rows_per_warp - 1) / rows_per_warp;   // all lanes stay converged for the shuffles
  for (long long it = warp; it < n_iter; it += n_warps) {
    const long long r = it * rows_per_warp + lane / G;
    const bool row_ok = r < n;
    const unsigned long long rid = (unsigned long long)(row_base + (row_ok ? r : 0));
    float f = scale;
    float zk[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // the lane's normals of pass 1, reused below when it owns a single block
    const bool one_block = n_blk <= G;         // d <= 128: no second Philox pass
    if (family == RWMPT_P_UNIFORM_RADIUS) {
      // pass 1: squared norm of the row's normal vector (uniform.py:48-73: z / ||z|| * R * u^(1/d))
      float n2 = 0.0f;
      unsigned uw;
      if (one_block && n_blk < G) {
        // the group has idle lanes (d = 20: 5 blocks on 8 lanes): the first of them draws the row's radius word in the same Philox
        // call the others draw their normals in -- one call per lane instead of two, same counters, same values
        const uint4 w = philox4x32_10(sub < n_blk ? (unsigned)sub : 0xffffffffu, 0x50524F50u, (unsigned)rid, (unsigned)(rid >> 32), k0, k1);
        box_muller<false>(w.x, w.y, zk[0], zk[1]);
        box_muller<false>(w.z, w.w, zk[2], zk[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (sub < n_blk && 4 * sub + q < d) n2 = fmaf(zk[q], zk[q], n2);
        for (int o = G >> 1; o > 0; o >>= 1) n2 += __shfl_xor_sync(kFull, n2, o);
        uw = __shfl_sync(kFull, w.x, (lane & ~(G - 1)) + n_blk);
      } else {
        for (int b = sub; b < n_blk; b += G) {
          const uint4 w = philox4x32_10((unsigned)b, 0x50524F50u, (unsigned)rid, (unsigned)(rid >> 32), k0, k1);
          float z[4];
          box_muller<false>(w.x, w.y, z[0], z[1]);
          box_muller<false>(w.z, w.w, z[2], z[3]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (4 * b + q < d) n2 = fmaf(z[q], z[q], n2);
            zk[q] = z[q];
          }
        }
        for (int o = G >> 1; o > 0; o >>= 1) n2 += __shfl_xor_sync(kFull, n2, o);
        uw = philox4x32_10(0xffffffffu, 0x50524F50u, (unsigned)rid, (unsigned)(rid >> 32), k0, k1).x;
      }
      const float nrm = sqrt_approx(n2);
      const float safe = nrm > 1e-12f ? nrm : 1.0f;
      f = scale * ex2_approx(lg2_approx(u01_from_bits(uw)) / (float)d) * rcp_approx(safe);
    }
    for (int b = sub; b < n_blk; b += G) {
      float v[4];
      if (family == RWMPT_P_UNIFORM_RADIUS && one_block) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = zk[q] * f;
      } else {
      const uint4 w = philox4x32_10((unsigned)b, 0x50524F50u, (unsigned)rid, (unsigned)(rid >> 32), k0, k1);
      if (family == RWMPT_P_LAPLACE) {
        // laplace.py:47-69, as in pair_transform: 23 word bits -> r = 2u in [-1, 1); |increment| = -s ln(max(1 - |r|, 1e-6))
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = 4 * b + q;
          const float ds = (dscale && i < d) ? dscale[i] : 1.0f;
          const float fbits = __uint_as_float((ww[q] & 0x007fffffu) | 0x3f800000u);
          const float rr = fmaf(2.0f, fbits, -3.0f);
          v[q] = copysignf(lg2_approx(fmaxf(1.0f - fabsf(rr), 1e-6f)) * (scale * ds * kLn2), rr);
        }
      } else {
        box_muller<false>(w.x, w.y, v[0], v[1]);
        box_muller<false>(w.z, w.w, v[2], v[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] *= f;
      }
      }
      if (row_ok) {
        if (vec4) {
          *reinterpret_cast<float4*>(out + r * d + 4 * b) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * b + q < d) out[r * d + 4 * b + q] = v[q];
        }
      }
    }
  }
}

// ---- proposal sampler, flat form for Normal / Laplace rows that are whole float4 blocks (d % 4 == 0, aligned output): the
// output is one array of n * d / 4 blocks and every thread turns one Philox call into one float4 -- no idle lanes whatever d
// is (the grouped form above leaves 3 of 8 lanes idle at d = 20).  Same counters as proposal_kernel: identical values.
__global__ void __launch_bounds__(256) proposal_flat_kernel(int family, int n_blk, float scale, const float* __restrict__ dscale,
                                                            long long n_blocks, unsigned k0, unsigned k1, long long row_base,
                                                            float4* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_blocks; q += stride) {
    const long long r = q / n_blk;
    const int b = (int)(q - r * n_blk);
    const unsigned long long rid = (unsigned long long)(row_base + r);
    const uint4 w = philox4x32_10((unsigned)b, 0x50524F50u, (unsigned)rid, (unsigned)(rid >> 32), k0, k1);
    float v[4];
    if (family == RWMPT_P_LAPLACE) {
      const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ds = dscale ? dscale[4 * b + j] : 1.0f;
        const float fbits = __uint_as_float((ww[j] & 0x007fffffu) | 0x3f800000u);
        const float rr = fmaf(2.0f, fbits, -3.0f);
        v[j] = copysignf(lg2_approx(fmaxf(1.0f - fabsf(rr), 1e-6f)) * (scale * ds * kLn2), rr);
      }
    } else {
      box_muller<false>(w.x, w.y, v[0], v[1]);
      box_muller<false>(w.z, w.w, v[2], v[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= scale;
    }
    out[q] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- stand-alone PT swap sweep: one CTA keeps a whole ladder in shared memory --------------------
__global__ void __launch_bounds__(128) pt_swap_kernel(float* __restrict__ state, float* __restrict__ logp,
                                                      const float* __restrict__ beta, long long n_ladders, int K, int d,
                                                      int swap_mode, const float* __restrict__ su, unsigned k0, unsigned k1,
                                                      long long ladder_base, long long round_index,
                                                      unsigned char* __restrict__ dec, unsigned long long* __restrict__ acc_out) {
  extern __shared__ float sm[];
  float* s_lp = sm;               // [K]
  float* s_b = sm + K;            // [K]
  int* s_src = (int*)(sm + 2 * K);  // [K]
  float* s_x = sm + 3 * K;        // [K, d]
  for (long long l = blockIdx.x; l < n_ladders; l += gridDim.x) {
    const long long c0 = l * K;
    for (int i = threadIdx.x; i < K * d; i += blockDim.x) s_x[i] = state[c0 * d + i];
    for (int j = threadIdx.x; j < K; j += blockDim.x) { s_lp[j] = logp[c0 + j]; s_b[j] = beta[c0 + j]; s_src[j] = j; }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long lg = (unsigned long long)(ladder_base + l);
      for (int j = 0; j < K - 1; ++j) {
        const float u = su ? su[l * (K - 1) + j] : swap_uniform(k0, k1, lg, (unsigned long long)round_index, j);
        bool ok;
        if (swap_mode == RWMPT_SWAP_REFERENCE) {
          // slot j+1 still holds its pre-sweep occupant when pair j is examined
          ok = swap_accept<true>(s_b[j], s_b[j + 1], s_lp[j], s_lp[j + 1], u);
          if (ok) s_src[j] = j + 1;
        } else {
          const int sa = s_src[j], sb = s_src[j + 1];
          ok = swap_accept<true>(s_b[j], s_b[j + 1], s_lp[sa], s_lp[sb], u);
          if (ok) { s_src[j] = sb; s_src[j + 1] = sa; }
        }
        if (dec) dec[l * (K - 1) + j] = ok ? 1 : 0;
        if (acc_out && ok) acc_out[l * (K - 1) + j] += 1ull;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * d; i += blockDim.x) {
      const int j = i / d, q = i - j * d;
      const int src = s_src[j];
      if (src != j) state[(c0 + j) * d + q] = s_x[src * d + q];
    }
    for (int j = threadIdx.x; j < K; j += blockDim.x)
      if (s_src[j] != j) logp[c0 + j] = s_lp[s_src[j]];
    __syncthreads();
  }
}

// ---- ESJD reduction, bandwidth form (no per-row output wanted): the chain's retained rows are one flat array f[n*d] and
// sum_m ||x_m - x_{m-1}||^2 = sum_{j >= d} (f[j] - f[j-d])^2, so every thread streams aligned VW-float vectors of f and of
// f shifted by one row (the shifted stream re-reads lines the first stream fetched d floats earlier: L1 / L2 hits, HBM
// sees each byte once), four vector pairs in flight per thread, no shuffle inside the loop.  Algorithmic bytes: 4 n d
// per chain.  HBM-read bound.
template <int VW> struct VecOf;
template <> struct VecOf<4> { using T = float4; };
template <> struct VecOf<2> { using T = float2; };
template <> struct VecOf<1> { using T = float; };
__device__ __forceinline__ float sqdiff(float4 a, float4 b) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
  return fmaf(x, x, fmaf(y, y, fmaf(z, z, w * w)));
}
__device__ __forceinline__ float sqdiff(float2 a, float2 b) {
  const float x = a.x - b.x, y = a.y - b.y;
  return fmaf(x, x, y * y);
}
__device__ __forceinline__ float sqdiff(float a, float b) { return (a - b) * (a - b); }

template <int VW>
__global__ void __launch_bounds__(256) esjd_flat_kernel(const float* __restrict__ samples, long long stride, long long first,
                                                        long long n, int d, int ctas_per_chain, double* __restrict__ out) {
  using V = typename VecOf<VW>::T;
  const long long chain = blockIdx.x / ctas_per_chain;
  const int part = blockIdx.x % ctas_per_chain;
  const V* prev = reinterpret_cast<const V*>(samples + (chain * stride + first) * d);
  const V* cur = reinterpret_cast<const V*>(samples + (chain * stride + first + 1) * d);
  const long long total = (n - 1) * d / VW;                       // vectors of jumps (d % VW == 0)
  const long long per = (total + ctas_per_chain - 1) / ctas_per_chain;
  const long long lo = part * per;
  const long long hi = lo + per < total ? lo + per : total;
  double acc = 0.0;
  long long v = lo + threadIdx.x;
  for (; v + 3 * 256 < hi; v += 4 * 256) {
    const V c0 = cur[v], c1 = cur[v + 256], c2 = cur[v + 512], c3 = cur[v + 768];
    const V p0 = prev[v], p1 = prev[v + 256], p2 = prev[v + 512], p3 = prev[v + 768];
    acc += (double)((sqdiff(c0, p0) + sqdiff(c1, p1)) + (sqdiff(c2, p2) + sqdiff(c3, p3)));
  }
  for (; v < hi; v += 256) acc += (double)sqdiff(cur[v], prev[v]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  __shared__ double s_acc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_acc[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_acc[w];
    atomicAdd(&out[chain], t / (double)(n - 1));
  }
}

// ---- ESJD reduction WITH the per-chain count of rows that moved, bandwidth form.  Same streams as esjd_flat_kernel (aligned
// VW-float vectors of a row and of its predecessor), but a warp owns whole rows: with vpr = d / VW vectors per row it maps
// R = 32 / vpr rows onto R * vpr lanes per pass (vpr > 32: one row over ceil(vpr / 32) passes), so "did this row move" is
// an OR over the row's lanes -- one ballot per pass, no shuffle, no shared memory -- and the count is a popcount of row
// masks.  A row moved iff ANY coordinate differs from its predecessor (bitwise compare: an increment may be too small to
// change the squared sum but still change the state).  Four passes in flight per warp.
__device__ __forceinline__ bool differs(float4 a, float4 b) { return a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w; }
__device__ __forceinline__ bool differs(float2 a, float2 b) { return a.x != b.x || a.y != b.y; }
__device__ __forceinline__ bool differs(float a, float b) { return a != b; }
template <int VW> __device__ __forceinline__ typename VecOf<VW>::T vec_zero();
template <> __device__ __forceinline__ float4 vec_zero<4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float2 vec_zero<2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ float vec_zero<1>() { return 0.f; }

template <int VW>
__global__ void __launch_bounds__(256) esjd_moved_kernel(const float* __restrict__ samples, long long stride, long long first,
                                                         long long n, int d, int ctas_per_chain, double* __restrict__ out,
                                                         unsigned long long* __restrict__ moved) {
  using V = typename VecOf<VW>::T;
  const long long chain = blockIdx.x / ctas_per_chain;
  const int part = blockIdx.x % ctas_per_chain;
  const V* base = reinterpret_cast<const V*>(samples + (chain * stride + first) * d);   // row m starts at base + m * vpr
  const int vpr = d / VW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long jumps = n - 1;                                   // jump j compares row j + 1 with row j
  const long long per = (jumps + ctas_per_chain - 1) / ctas_per_chain;
  const long long lo = part * per, hi = lo + per < jumps ? lo + per : jumps;
  double acc = 0.0;
  unsigned long long mv = 0;
  if (vpr <= 32) {
    const int R = 32 / vpr;                                        // rows per pass
    const int r = lane / vpr, vec = lane - r * vpr;
    const bool lane_on = lane < R * vpr;
    const unsigned row_mask = vpr == 32 ? 0xffffffffu : ((1u << vpr) - 1u);
    constexpr int U = 4;                                           // passes in flight (8 and a loop-free row count measured slower)
    for (long long j0 = lo + (long long)warp * R * U; j0 < hi; j0 += (long long)n_warps * R * U) {
      V c[U], p[U];
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long j = j0 + (long long)u * R + r;
        on[u] = lane_on && j < hi;
        c[u] = on[u] ? base[(j + 1) * vpr + vec] : vec_zero<VW>();
        p[u] = on[u] ? base[j * vpr + vec] : vec_zero<VW>();
      }
      float s = 0.0f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s += sqdiff(c[u], p[u]);
        const unsigned b = __ballot_sync(kFull, on[u] && differs(c[u], p[u]));
        for (int q = 0; q < R; ++q) mv += ((b >> (q * vpr)) & row_mask) != 0u;   // uniform across the warp
      }
      acc += (double)s;
    }
    if (lane != 0) mv = 0;                                          // every lane counted the same rows
  } else {
    const int passes = (vpr + 31) / 32;
    for (long long j = lo + warp; j < hi; j += n_warps) {
      float s = 0.0f;
      unsigned any = 0u;
      for (int q = 0; q < passes; ++q) {
        const int vec = q * 32 + lane;
        const bool on = vec < vpr;
        const V c = on ? base[(j + 1) * vpr + vec] : vec_zero<VW>();
        const V p = on ? base[j * vpr + vec] : vec_zero<VW>();
        s += sqdiff(c, p);
        any |= __ballot_sync(kFull, on && differs(c, p));
      }
      acc += (double)s;
      mv += (lane == 0 && any != 0u) ? 1ull : 0ull;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(kFull, acc, o);
    mv += __shfl_xor_sync(kFull, mv, o);
  }
  __shared__ double s_acc[8];
  __shared__ unsigned long long s_mv[8];
  if (lane == 0) { s_acc[warp] = acc; s_mv[warp] = mv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    unsigned long long tm = 0;
    for (int w = 0; w < n_warps; ++w) { t += s_acc[w]; tm += s_mv[w]; }
    atomicAdd(&out[chain], t / (double)(n - 1));
    atomicAdd(&moved[chain], tm);
  }
}

// ---- peak probes: dependent-free FFMA and MUFU.EX2 loops, 8 independent chains per thread ---------
__global__ void __launch_bounds__(256) probe_ffma_kernel(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1e-3f * (threadIdx.x + i);
  const float b = 0.999f, c = 1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_mufu_kernel(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1e-3f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2_approx(-a[i]);
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}

__global__ void philox_kat_kernel(const uint32_t* in, uint32_t* out) {
  const uint4 r = philox4x32_10(in[0], in[1], in[2], in[3], in[4], in[5]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // namespace rwmpt

using namespace rwmpt;

extern "C" {

int rwmpt_version(void) { return RWMPT_VERSION; }
const char* rwmpt_last_error(void) { return g_err; }
uint64_t rwmpt_sizeof_run_args(void) { return sizeof(rwmpt_run_args_t); }

int rwmpt_rwm_run(const rwmpt_run_args_t* args, void* cuda_stream) { return run_impl(args, cuda_stream, true); }
int rwmpt_pt_run(const rwmpt_run_args_t* args, void* cuda_stream) { return run_impl(args, cuda_stream, false); }

int64_t rwmpt_count_swap_rounds(int64_t step_offset, int64_t n_steps, int64_t burn_in, int32_t swap_every) {
  return count_rounds(step_offset, n_steps, burn_in, swap_every);
}

int rwmpt_pick_lanes(int32_t dim, int32_t n_temps, int64_t n_ladders, int32_t math_mode, int32_t* elems_per_lane) {
  LaunchGeom g;
  if (dim < 1 || n_temps < 1 || n_ladders < 1) return fail(RWMPT_EINVAL, "dim, n_temps, n_ladders must be >= 1");
  const int rc = pick_geometry(dim, n_temps, n_ladders, math_mode == RWMPT_MATH_IEEE, 0, &g);
  if (rc) return rc;
  if (elems_per_lane) *elems_per_lane = g.E;
  return g.W;
}

int rwmpt_pick_geometry(const rwmpt_run_args_t* r, int32_t* lanes_per_chain, int32_t* elems_per_lane) {
  if (!r) return fail(RWMPT_EINVAL, "args is NULL");
  if (r->target.dim < 1 || r->n_temps < 1 || r->n_ladders < 1) return fail(RWMPT_EINVAL, "dim, n_temps, n_ladders must be >= 1");
  LaunchGeom g;
  const int rc = pick_geometry(r->target.dim, r->n_temps, r->n_ladders, r->math_mode == RWMPT_MATH_IEEE, r->lanes_per_chain, &g,
                               r->target.family, r->proposal_family, true);
  if (rc) return rc;
  if (lanes_per_chain) *lanes_per_chain = g.W;
  if (elems_per_lane) *elems_per_lane = g.E;
  return RWMPT_OK;
}

int rwmpt_log_density(const rwmpt_target_t* target, const float* x, int64_t n, float* out, int32_t math_mode,
                      void* cuda_stream) {
  int rc = check_target(target);
  if (rc) return rc;
  if (n < 0) return fail(RWMPT_EINVAL, "n must be >= 0");
  if (n == 0) return RWMPT_OK;
  if (!x || !out) return fail(RWMPT_EINVAL, "x and out must be non-NULL");
  LaunchGeom g;
  const bool ieee = math_mode == RWMPT_MATH_IEEE;
  rc = pick_geometry(target->dim, 1, n, ieee, 0, &g);
  if (rc) return rc;
  cudaError_t e = dispatch_logp(target->family, target->params, target->dim, g.E, g.W, x, n, out, ieee, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "log-density kernel launch");
  return RWMPT_OK;
}

int rwmpt_swap_prob_estimate(const rwmpt_target_t* target, float beta_curr, float beta_star, int64_t n, uint64_t seed,
                             int64_t row_id_base, double* sum_out, void* cuda_stream) {
  int rc = check_target(target);
  if (rc) return rc;
  if (n < 0) return fail(RWMPT_EINVAL, "n must be >= 0");
  if (!(beta_curr > 0.0f) || !(beta_star > 0.0f)) return fail(RWMPT_EINVAL, "beta_curr and beta_star must be positive");
  if (!sum_out) return fail(RWMPT_EINVAL, "sum_out is NULL");
  if (n == 0) return RWMPT_OK;
  LaunchGeom g;
  rc = pick_geometry(target->dim, 1, n, false, 0, &g);
  if (rc) return rc;
  cudaError_t e = dispatch_swap_prob(target->family, target->params, target->dim, g.E, g.W, beta_curr, beta_star, n,
                                     (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32), row_id_base, sum_out,
                                     (cudaStream_t)cuda_stream);
  if (e == cudaErrorNotSupported) {
    cudaGetLastError();
    return fail(RWMPT_ENOTSUP, "target family %d has no native tempered sampler (use draw_samples_torch + rwmpt_log_density)",
                target->family);
  }
  if (e != cudaSuccess) return cuda_fail(e, "swap-probability kernel launch");
  return RWMPT_OK;
}

int rwmpt_proposal_sample(int32_t proposal_family, int32_t dim, float scale, const float* dim_scale, int64_t n,
                          uint64_t seed, int64_t row_id_base, float* out, void* cuda_stream) {
  if (proposal_family < RWMPT_P_NORMAL || proposal_family > RWMPT_P_UNIFORM_RADIUS)
    return fail(RWMPT_EINVAL, "unknown proposal family %d", proposal_family);
  if (dim < 1 || n < 0) return fail(RWMPT_EINVAL, "dim must be >= 1 and n >= 0");
  if (!(scale > 0.0f)) return fail(RWMPT_EINVAL, "scale must be positive");
  if (n == 0) return RWMPT_OK;
  if (!out) return fail(RWMPT_EINVAL, "out is NULL");
  const int n_blk = (dim + 3) / 4;
  int G = 1;
  while (G < n_blk && G < 32) G <<= 1;
  const long long rows_per_cta = 4LL * (32 / G);
  long long blocks = (n + rows_per_cta - 1) / rows_per_cta;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int vec4 = (dim % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) ? 1 : 0;
  if (vec4 && proposal_family != RWMPT_P_UNIFORM_RADIUS) {
    const long long n_blocks = n * n_blk;
    long long ctas = (n_blocks + 255) / 256;
    if (ctas > 148 * 16) ctas = 148 * 16;
    proposal_flat_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)cuda_stream>>>(proposal_family, n_blk, scale, dim_scale, n_blocks,
                                                                             (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32),
                                                                             row_id_base, reinterpret_cast<float4*>(out));
    cudaError_t ef = cudaGetLastError();
    if (ef != cudaSuccess) return cuda_fail(ef, "proposal kernel launch");
    return RWMPT_OK;
  }
  proposal_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)cuda_stream>>>(proposal_family, dim, G, scale, dim_scale, n,
                                                                         (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32),
                                                                         row_id_base, vec4, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "proposal kernel launch");
  return RWMPT_OK;
}

int rwmpt_pt_swap(float* state, float* logp, const float* beta, int64_t n_ladders, int32_t n_temps, int32_t dim,
                  int32_t swap_mode, const float* swap_uniforms, uint64_t seed, int64_t ladder_id_base,
                  int64_t round_index, unsigned char* swap_decisions, unsigned long long* swap_accepts,
                  void* cuda_stream) {
  if (!state || !logp || !beta) return fail(RWMPT_EINVAL, "state, logp, beta must be non-NULL");
  if (n_temps < 1 || dim < 1 || n_ladders < 0) return fail(RWMPT_EINVAL, "bad n_temps / dim / n_ladders");
  if (swap_mode != RWMPT_SWAP_REFERENCE && swap_mode != RWMPT_SWAP_EXCHANGE) return fail(RWMPT_EINVAL, "bad swap_mode");
  if (n_ladders == 0 || n_temps == 1) return RWMPT_OK;
  const size_t smem = ((size_t)3 * n_temps + (size_t)n_temps * dim) * sizeof(float);
  if (smem > 200 * 1024) return fail(RWMPT_ENOTSUP, "ladder of %d x %d floats does not fit in shared memory", n_temps, dim);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pt_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(pt_swap_kernel)");
  }
  long long blocks = n_ladders < 148 * 8 ? n_ladders : 148 * 8;
  pt_swap_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)cuda_stream>>>(
      state, logp, beta, n_ladders, n_temps, dim, swap_mode, swap_uniforms, (unsigned)(seed & 0xffffffffu),
      (unsigned)(seed >> 32), ladder_id_base, round_index, swap_decisions, swap_accepts);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pt_swap kernel launch");
  return RWMPT_OK;
}

int rwmpt_esjd_reduce(const float* samples, int64_t n_chains, int64_t stride, int64_t first, int64_t n, int32_t dim,
                      double* esjd_out, unsigned long long* moved_out, void* cuda_stream) {
  if (!samples || !esjd_out) return fail(RWMPT_EINVAL, "samples and esjd_out must be non-NULL");
  if (n_chains < 0 || dim < 1 || first < 0 || n < 0 || first + n > stride) return fail(RWMPT_EINVAL, "bad sizes for esjd_reduce");
  if (n_chains == 0) return RWMPT_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaError_t e = cudaMemsetAsync(esjd_out, 0, sizeof(double) * n_chains, st);
  if (e != cudaSuccess) return cuda_fail(e, "memset esjd_out");
  if (moved_out) {
    e = cudaMemsetAsync(moved_out, 0, sizeof(unsigned long long) * n_chains, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset moved_out");
  }
  if (n < 2) return RWMPT_OK;  // fewer than two rows: ESJD is 0 (rwm_gpu_optimized.py:526-527)
  if (!moved_out) {
    // bandwidth form: 148 SMs x 8 resident CTAs, each with at least ~16 KiB of jumps to stream
    const uintptr_t p = reinterpret_cast<uintptr_t>(samples);
    const int vw = (dim % 4 == 0 && p % 16 == 0) ? 4 : ((dim % 2 == 0 && p % 8 == 0) ? 2 : 1);
    const long long vecs = (n - 1) * dim / vw;
    long long want = (148LL * 8 + n_chains - 1) / n_chains;
    long long maxp = (vecs + 1023) / 1024;
    long long cpc = want < 1 ? 1 : (want > maxp ? maxp : want);
    if (cpc < 1) cpc = 1;
    if (n_chains * cpc > 2147483647LL) return fail(RWMPT_ENOTSUP, "too many chains for esjd_reduce");
    const unsigned grid = (unsigned)(n_chains * cpc);
    if (vw == 4) esjd_flat_kernel<4><<<grid, 256, 0, st>>>(samples, stride, first, n, dim, (int)cpc, esjd_out);
    else if (vw == 2) esjd_flat_kernel<2><<<grid, 256, 0, st>>>(samples, stride, first, n, dim, (int)cpc, esjd_out);
    else esjd_flat_kernel<1><<<grid, 256, 0, st>>>(samples, stride, first, n, dim, (int)cpc, esjd_out);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "esjd kernel launch");
    return RWMPT_OK;
  }
  long long want = (148LL * 8 + n_chains - 1) / n_chains;  // enough CTAs to fill the machine
  long long maxp = (n - 1 + 255) / 256;                    // at least ~256 jumps per CTA
  int cpc = (int)(want < 1 ? 1 : (want > maxp ? maxp : want));
  if (cpc < 1) cpc = 1;
  if (n_chains * cpc > 2147483647LL) return fail(RWMPT_ENOTSUP, "too many chains for esjd_reduce");
  {
    const uintptr_t p = reinterpret_cast<uintptr_t>(samples);
    const int vw = (dim % 4 == 0 && p % 16 == 0) ? 4 : ((dim % 2 == 0 && p % 8 == 0) ? 2 : 1);
    const unsigned grid = (unsigned)(n_chains * cpc);
    if (vw == 4) esjd_moved_kernel<4><<<grid, 256, 0, st>>>(samples, stride, first, n, dim, cpc, esjd_out, moved_out);
    else if (vw == 2) esjd_moved_kernel<2><<<grid, 256, 0, st>>>(samples, stride, first, n, dim, cpc, esjd_out, moved_out);
    else esjd_moved_kernel<1><<<grid, 256, 0, st>>>(samples, stride, first, n, dim, cpc, esjd_out, moved_out);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "esjd kernel launch");
  return RWMPT_OK;
}

int rwmpt_probe_peaks(double* fp32_tflops, double* sfu_gops) {
  float* d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, sizeof(float));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  const int blocks = 148 * 8, threads = 256;
  double best[2] = {0.0, 0.0};
  for (int kind = 0; kind < 2; ++kind) {
    const int iters = kind == 0 ? 4096 : 1024;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(t0);
      if (kind == 0) probe_ffma_kernel<<<blocks, threads>>>(d_out, iters);
      else probe_mufu_kernel<<<blocks, threads>>>(d_out, iters);
      cudaEventRecord(t1);
      e = cudaEventSynchronize(t1);
      if (e != cudaSuccess) break;
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, t0, t1);
      const double ops = (double)blocks * threads * (double)iters * 64.0;
      const double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best[kind]) best[kind] = rate;
    }
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "peak probe");
  if (fp32_tflops) *fp32_tflops = 2.0 * best[0] / 1e12;
  if (sfu_gops) *sfu_gops = best[1] / 1e9;
  return RWMPT_OK;
}

// issue-rate probe at a chosen occupancy: kind 0 = FFMA (8 independent chains), 1 = MUFU.EX2, 2 = Philox rounds
__global__ void probe_philox_kernel(uint32_t* out, int iters) {
  uint32_t c0 = threadIdx.x, c1 = blockIdx.x, c2 = 7, c3 = 9, d0 = threadIdx.x * 3, d1 = 1, d2 = 2, d3 = 3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      philox_round(c0, c1, c2, c3, 0x1234567u + r, 0x89abcdefu + r);
      philox_round(d0, d1, d2, d3, 0x1234567u + r, 0x89abcdefu + r);
    }
  }
  if ((c0 ^ c1 ^ c2 ^ c3 ^ d0 ^ d1 ^ d2 ^ d3) == 0x12345u) out[0] = c0;
}

int rwmpt_probe_issue(int kind, int blocks, int threads, int iters, double* ops_per_s) {
  float* d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, sizeof(float));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(t0);
    if (kind == 0) probe_ffma_kernel<<<blocks, threads>>>(d_out, iters);
    else if (kind == 1) probe_mufu_kernel<<<blocks, threads>>>(d_out, iters);
    else probe_philox_kernel<<<blocks, threads>>>((uint32_t*)d_out, iters);
    cudaEventRecord(t1);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, t0, t1);
    // warp-instructions issued: kinds 0/1: 64 per iteration per thread; kind 2: 16 rounds x 6 instr (2 IMAD.HI, 2 IMAD, 2 LOP3)
    const double per_iter = kind == 2 ? 96.0 : 64.0;
    const double winstr = (double)blocks * (threads / 32.0) * iters * per_iter;
    const double rate = winstr / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "issue probe");
  if (ops_per_s) *ops_per_s = best;
  return RWMPT_OK;
}

int rwmpt_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t h[6] = {ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]};
  uint32_t *d_in = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_in, sizeof(h));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  e = cudaMalloc(&d_out, 4 * sizeof(uint32_t));
  if (e != cudaSuccess) { cudaFree(d_in); return cuda_fail(e, "cudaMalloc"); }
  cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
  philox_kat_kernel<<<1, 1>>>(d_in, d_out);
  e = cudaMemcpy(out, d_out, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "philox KAT");
  return RWMPT_OK;
}

// ---- host-buffer end-to-end entry ---------------------------------------------------------------
namespace {
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void* host = nullptr;
  bool out = false;
  size_t pitch = 0, row_off = 0, width = 0;  // pitch != 0: copied back as a 2-D block of rows (retained samples)
};
}  // namespace

int rwmpt_run_host(const rwmpt_run_args_t* r, int32_t device, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!r) return fail(RWMPT_EINVAL, "args is NULL");
  int prev_device = -1;
  cudaGetDevice(&prev_device);  // restored before returning: the entry must not change the caller's current device
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  struct RestoreDevice {
    int dev;
    ~RestoreDevice() { if (dev >= 0) cudaSetDevice(dev); }
  } restore{prev_device};
  const int64_t d = r->target.dim, K = r->n_temps;
  if (d < 1 || K < 1 || r->n_ladders < 0 || r->n_steps < 0) return fail(RWMPT_EINVAL, "bad sizes");
  const int64_t n_chains = r->n_ladders * K;
  const int64_t rounds = K > 1 ? count_rounds(r->step_offset, r->n_steps, r->burn_in, r->swap_every) : 0;
  const int64_t stored_chains = r->store_mode == RWMPT_STORE_ALL ? n_chains : r->n_ladders;
  rwmpt_run_args_t a = *r;
  std::vector<DevBuf> bufs;
  uint64_t h2d = 0, d2h = 0;
  cudaStream_t st;
  e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  int rc = RWMPT_OK;
  // Retained-sample rows THIS call writes: global step s is retained as row (s - store_start) / thin - 1 when
  // s > store_start and (s - store_start) % thin == 0; this call runs steps (step_offset, step_offset + n_steps].  Only
  // those rows are copied back (one 2-D copy per buffer: `stored_chains` segments, one per chain), so the rows earlier
  // calls of a resumed run returned -- and rows at or beyond sample_rows -- are never overwritten on the host with
  // device memory this call did not write.
  int64_t row_lo = 0, row_hi = 0;
  if (r->samples && r->thin >= 1) {
    const int64_t before = r->step_offset > r->store_start ? (r->step_offset - r->store_start) / r->thin : 0;
    const int64_t upto = r->step_offset + r->n_steps > r->store_start ? (r->step_offset + r->n_steps - r->store_start) / r->thin : 0;
    row_lo = before < r->sample_rows ? before : r->sample_rows;
    row_hi = upto < r->sample_rows ? upto : r->sample_rows;
  }
  // one device arena for every staged buffer (a single stream-ordered allocation per call), 256-byte aligned slices
  struct Want { const void* host; size_t bytes; bool in, out; void** slot; size_t pitch, row_off, width; };
  std::vector<Want> wants;
  auto stage = [&](const void* host, size_t bytes, bool in, bool out, void** slot) {
    *slot = nullptr;
    if (host && bytes) wants.push_back({host, bytes, in, out, slot, 0, 0, 0});
  };
  // out-only, copied back as `stored_chains` segments of rows [row_lo, row_hi): pitch / offset / width in bytes
  auto stage_rows = [&](const void* host, size_t row_bytes, void** slot) {
    *slot = nullptr;
    const size_t bytes = row_bytes * (size_t)stored_chains * (size_t)r->sample_stride;
    if (host && bytes)
      wants.push_back({host, bytes, false, row_hi > row_lo, slot, row_bytes * (size_t)r->sample_stride, row_bytes * (size_t)row_lo,
                       row_bytes * (size_t)(row_hi - row_lo)});
  };
  stage(r->target.params, sizeof(float) * r->target.n_params, true, false, (void**)&a.target.params);
  stage(r->prop_scale, sizeof(float) * n_chains, true, false, (void**)&a.prop_scale);
  stage(r->prop_dim_scale, sizeof(float) * d, true, false, (void**)&a.prop_dim_scale);
  stage(r->beta, sizeof(float) * n_chains, true, false, (void**)&a.beta);
  stage(r->state, sizeof(float) * n_chains * d, true, true, (void**)&a.state);
  stage(r->logp, sizeof(float) * n_chains, true, true, (void**)&a.logp);
  stage_rows(r->samples, sizeof(float) * d, (void**)&a.samples);
  stage_rows(r->sample_logp, sizeof(float), (void**)&a.sample_logp);
  stage(r->accept_count, 8 * n_chains, true, true, (void**)&a.accept_count);
  stage(r->sq_jump_sum, 8 * n_chains, true, true, (void**)&a.sq_jump_sum);
  stage(r->swap_accepts, 8 * r->n_ladders * (K > 1 ? K - 1 : 0), true, true, (void**)&a.swap_accepts);
  stage(r->swap_last_attempt, 8 * n_chains, true, true, (void**)&a.swap_last_attempt);
  stage(r->inj_increments, sizeof(float) * r->n_steps * n_chains * d, true, false, (void**)&a.inj_increments);
  stage(r->inj_uniforms, sizeof(float) * r->n_steps * n_chains, true, false, (void**)&a.inj_uniforms);
  stage(r->inj_swap_uniforms, sizeof(float) * rounds * r->n_ladders * (K - 1), true, false, (void**)&a.inj_swap_uniforms);
  // decision outputs: the kernel writes every element of both (one flag per step and chain / per sweep and pair)
  stage(r->decisions, (size_t)r->n_steps * n_chains, false, true, (void**)&a.decisions);
  stage(r->swap_decisions, (size_t)rounds * r->n_ladders * (K > 1 ? K - 1 : 0), false, true, (void**)&a.swap_decisions);
  size_t total = 0;
  for (auto& w : wants) total += (w.bytes + 255) & ~(size_t)255;
  char* arena = nullptr;
  if (total) {
    // stream-ordered allocation from the device's default pool, which keeps its memory between calls (release threshold
    // raised once per device): after the first call this is a pool hit, not a cudaMalloc / cudaFree pair that
    // serialises against every other process and context on the node
    static std::atomic<unsigned long long> pool_kept{0ull};   // bit per device; the only process-wide state of the library
    if (device >= 0 && device < 64) {
      const unsigned long long bit = 1ull << device;
      if (!(pool_kept.fetch_or(bit) & bit)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
          unsigned long long keep = ~0ull;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
      }
    }
    e = cudaMallocAsync((void**)&arena, total, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync");
  }
  size_t off = 0;
  for (auto& w : wants) {
    if (rc) break;
    DevBuf b;
    b.p = arena + off; b.bytes = w.bytes; b.host = const_cast<void*>(w.host); b.out = w.out;
    b.pitch = w.pitch; b.row_off = w.row_off; b.width = w.width;
    off += (w.bytes + 255) & ~(size_t)255;
    *w.slot = b.p;
    if (w.in) {
      e = cudaMemcpyAsync(b.p, w.host, w.bytes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) rc = cuda_fail(e, "H2D copy");
      h2d += w.bytes;
    }
    bufs.push_back(b);
  }
  if (!rc) rc = run_impl(&a, st, false);
  if (!rc) {
    for (auto& b : bufs) {
      if (!b.out) continue;
      if (b.pitch) {
        e = cudaMemcpy2DAsync((char*)b.host + b.row_off, b.pitch, (char*)b.p + b.row_off, b.pitch, b.width, (size_t)stored_chains,
                              cudaMemcpyDeviceToHost, st);
        d2h += b.width * (size_t)stored_chains;
      } else {
        e = cudaMemcpyAsync(b.host, b.p, b.bytes, cudaMemcpyDeviceToHost, st);
        d2h += b.bytes;
      }
      if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy"); break; }
    }
  }
  if (arena) cudaFreeAsync(arena, st);
  e = cudaStreamSynchronize(st);
  if (!rc && e != cudaSuccess) rc = cuda_fail(e, "stream synchronize");
  cudaStreamDestroy(st);
  if (h2d_bytes) *h2d_bytes = h2d;
  if (d2h_bytes) *d2h_bytes = d2h;
  return rc;
}

}  // extern "C"
