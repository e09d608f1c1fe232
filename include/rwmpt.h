/*
 * rwmpt.h -- C ABI of librwmpt.so: the B200 (sm_100a) Random-Walk-Metropolis / Parallel-Tempering
 * sampling hot path.
 *
 * The reference (aidanmrli/rwm-pt-pytorch) has no FFI of its own: its boundary for this path is the
 * Python class protocol.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference repository root).  The Python facade classes in
 * rwm_pt_pytorch_b200/ (same names and arguments as the reference's) bind these symbols with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - unless a function name ends in `_host`, every data pointer is a DEVICE pointer owned by the
 *     caller, the work is enqueued on `cuda_stream` (a cudaStream_t cast to void*; NULL = default
 *     stream) and the call returns without synchronising;
 *   - return value 0 = ok, negative = error (RWMPT_E*); rwmpt_last_error() gives a thread-local text;
 *   - re-entrant, any number of streams / devices; the only process-wide state is one atomic bit per device that
 *     remembers rwmpt_run_host has raised the device memory pool's release threshold;
 *   - all floating point data is IEEE binary32 unless stated.
 */
#ifndef RWMPT_H_
#define RWMPT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RWMPT_VERSION 100 /* major*100 + minor */

#define RWMPT_OK 0
#define RWMPT_EINVAL (-1)  /* bad argument (the Python facade raises ValueError)          */
#define RWMPT_ENOTSUP (-2) /* valid but unsupported shape/family (NotImplementedError)    */
#define RWMPT_ECUDA (-3)   /* CUDA runtime / launch error (RuntimeError)                  */

/* Target families: the `log_density` methods of target_distributions/*_torch.py.
 * `params` is a device array of floats: 16 scalars followed by per-dimension vectors.          */
enum {
  RWMPT_T_ROUGH_CARPET = 0,      /* multimodal_torch.py:470-510   P[0..2] modes, P[3..5] log w, P[6] log sqrt(2pi),
                                    P[7] has_scaling, P[8] sum log s; P[16..16+d) s                              */
  RWMPT_T_THREE_MIXTURE = 1,     /* multimodal_torch.py:173-242   P[0..2] log w, P[3..5] c1_k, P[6] scaled, P[7] log|J|;
                                    P[16..16+3d) means (k-major), then d scaling factors                         */
  RWMPT_T_FULL_ROSENBROCK = 2,   /* rosenbrock_torch.py:67-84     P[0] a, P[1] b; P[16..16+d-1) mu               */
  RWMPT_T_EVEN_ROSENBROCK = 3,   /* rosenbrock_torch.py:194-210   P[0] a, P[1] b; P[16..16+d/2) mu               */
  RWMPT_T_HYBRID_ROSENBROCK = 4, /* rosenbrock_torch.py:312-351   P[0] a, P[1] b, P[2] mu, P[3] n1, P[4] n2      */
  RWMPT_T_NEAL_FUNNEL = 5,       /* funnel_torch.py:39-76         P[0] mu_v, P[1] sigma_v^2, P[2] mu_z,
                                    P[3] log sigma_v^2, P[4] log 2pi, P[5] d-1                                   */
  RWMPT_T_HYPERCUBE = 6,         /* hypercube_torch.py:49-80      P[0] L, P[1] R, P[2] log density               */
  RWMPT_T_IID_GAMMA = 7,         /* iid_product_torch.py:52-91    P[0] shape, P[1] scale, P[2] log norm const    */
  RWMPT_T_IID_BETA = 8,          /* iid_product_torch.py:188-229  P[0] alpha, P[1] beta, P[2] log norm const     */
  RWMPT_T_SCALED_MVN = 9,        /* multivariate_normal_torch.py:198-223  P[0] log norm const; P[16..16+d) c     */
  RWMPT_T_MVN_DIAG = 10,         /* multivariate_normal_torch.py:62-92 with diagonal cov: P[0] log norm const;
                                    P[16..16+d) mean, then d diagonal precisions                                 */
  RWMPT_T_MVN_DENSE = 11,        /* multivariate_normal_torch.py:62-92 with a general covariance: P[0] log norm const;
                                    P[16..16+d) mean, then cov_inv row-major (d x d); dim <= 128                  */
  RWMPT_T_SUPER_FUNNEL = 12,     /* funnel_torch.py:193-291 (hierarchical logistic regression)  P[0] J, P[1] K,
                                    P[2] prior hyper-mean var, P[3] its log, P[4] prior tau scale, P[5] its log,
                                    P[6] log 2pi, P[7] log 2, P[8] log pi, P[9] N; then N records (group, y, x[K]);
                                    dim = J + J K + K + 3 <= 128                                                 */
  RWMPT_T_COUNT = 13
};
#define RWMPT_PARAM_HEADER 16

/* Proposal families: proposal_distributions/{normal,laplace,uniform}.py */
enum { RWMPT_P_NORMAL = 0, RWMPT_P_LAPLACE = 1, RWMPT_P_UNIFORM_RADIUS = 2 };
/* Swap semantics: 0 = what the reference does (accepted pair copies k -> j, k unchanged;
 * pt_rwm_gpu_optimized.py:50-59), 1 = textbook exchange (pt_rwm.py:137-155).                   */
enum { RWMPT_SWAP_REFERENCE = 0, RWMPT_SWAP_EXCHANGE = 1 };
/* Arithmetic: 0 = fast intrinsics (MUFU ex2/lg2/sin/cos, FMA contraction), 1 = IEEE, un-fused,
 * in the reference's operation order (SURVEY.md section 8a') -- the parity mode.                          */
enum { RWMPT_MATH_FAST = 0, RWMPT_MATH_IEEE = 1 };
/* Trajectory storage: nothing / temperature 0 of every ladder / every chain.                   */
enum { RWMPT_STORE_NONE = 0, RWMPT_STORE_COLD = 1, RWMPT_STORE_ALL = 2 };
/* AUTO: balanced when a plain launch would load the SMs' warp schedulers unevenly (few warps per scheduler);
   PLAIN: one CTA per unit for the whole run; BALANCED: SM-sized grid, time-sliced units handed out by ticket. */
enum { RWMPT_SCHEDULE_AUTO = 0, RWMPT_SCHEDULE_PLAIN = 1, RWMPT_SCHEDULE_BALANCED = 2, RWMPT_SCHEDULE_SPECIALISED = 3 };

typedef struct {
  int32_t family;      /* RWMPT_T_*                                  */
  int32_t dim;         /* d >= 1                                     */
  const float* params; /* device, n_params floats, layout above      */
  int64_t n_params;
} rwmpt_target_t;

/*
 * One run of `n_steps` Metropolis steps for n_ladders * n_temps chains (RWM: n_temps = 1).
 * Chain c = ladder * n_temps + temperature index.  Replaces the Python hot loops
 *   RandomWalkMH_GPU_Optimized.generate_samples / _single_step_ultra_fused
 *       (algorithms/rwm_gpu_optimized.py:402-488, 289-336, 9-32) and
 *   ParallelTemperingRWM_GPU_Optimized.generate_samples / step / _attempt_all_swaps / _add_states_to_chains
 *       (algorithms/pt_rwm_gpu_optimized.py:694-770, 541-574, 594-633, 635-653),
 * including the proposal plugins' sample() (proposal_distributions/*.py) and target.log_density().
 * Global step index of local step t is s = step_offset + t + 1 (1-based like the reference's
 * `step_counter`); acceptances / jumps are accumulated for s > burn_in; a swap sweep runs after the
 * Metropolis step when s % swap_every == 0 and s > burn_in.  state/logp and every accumulator are
 * in/out, so a run can be resumed by calling again with step_offset advanced.
 */
typedef struct {
  rwmpt_target_t target;
  int32_t proposal_family;     /* RWMPT_P_*                                                              */
  int32_t n_temps;             /* K >= 1                                                                 */
  const float* prop_scale;     /* [n_chains] per-chain scale: Normal std sqrt(var/beta) (normal.py:27-31,
                                  pt_rwm_gpu_optimized.py:453-455), Laplace multiplier 1/sqrt(beta), ball
                                  radius r/sqrt(beta) (uniform.py:28-32)                                 */
  const float* prop_dim_scale; /* [dim] or NULL: per-dimension factor (Laplace sqrt(var_i/2), laplace.py:29-32) */
  const float* beta;           /* [n_chains] inverse temperatures                                        */
  int64_t n_ladders;
  int64_t n_steps;
  int64_t burn_in;
  int64_t step_offset;
  int32_t swap_every;          /* >= 1 (ignored when n_temps == 1)                                       */
  int32_t swap_mode;           /* RWMPT_SWAP_*                                                           */
  float* state;                /* [n_chains, dim] in/out                                                 */
  float* logp;                 /* [n_chains] in/out: log-density of `state` (see rwmpt_log_density)      */
  uint64_t seed;               /* Philox4x32-10 key                                                      */
  int64_t chain_id_base;       /* global id of local chain 0: Philox subsequence => GPU-count invariant  */
  /* retained samples (nullable) */
  float* samples;              /* [n_stored_chains, sample_stride, dim]; row m of a chain holds the state
                                  after global step s = store_start + thin*(m+1)                         */
  float* sample_logp;          /* [n_stored_chains, sample_stride] or NULL                               */
  int32_t store_mode;          /* RWMPT_STORE_*                                                          */
  int32_t math_mode;           /* RWMPT_MATH_*                                                           */
  int64_t store_start;
  int64_t thin;                /* >= 1                                                                   */
  int64_t sample_stride;       /* row stride between consecutive chains in `samples` / `sample_logp`     */
  int64_t sample_rows;         /* rows m >= sample_rows are silently dropped (the reference stops storing
                                  when its pre-allocated chain is full, pt_rwm_gpu_optimized.py:640)     */
  /* accumulators, all in/out (+=), nullable */
  unsigned long long* accept_count;        /* [n_chains] Metropolis acceptances, s > burn_in             */
  double* sq_jump_sum;                     /* [n_chains] sum ||x_s - x_{s-1}||^2 incl. swap moves, s > burn_in */
  unsigned long long* swap_accepts;        /* [n_ladders, n_temps-1] accepted swaps per adjacent pair    */
  unsigned long long* swap_last_attempt;   /* [n_chains] attempt counter value at this pair's last accepted
                                              swap (the reference refreshes swap_acceptance_rate / pt_esjd
                                              only then, pt_rwm_gpu_optimized.py:627-633); max-updated   */
  /* test mode: inject the reference's randomness (nullable) */
  const float* inj_increments;     /* [n_steps, n_chains, dim] increments AFTER scaling                  */
  const float* inj_uniforms;       /* [n_steps, n_chains]                                                */
  const float* inj_swap_uniforms;  /* [n_rounds, n_ladders, n_temps-1]                                   */
  unsigned char* decisions;        /* [n_steps, n_chains] out, nullable                                  */
  unsigned char* swap_decisions;   /* [n_rounds, n_ladders, n_temps-1] out, nullable                     */
  int32_t lanes_per_chain;         /* 0 = auto; else power of two <= 32: threads cooperating on a chain  */
  int32_t schedule;                /* RWMPT_SCHEDULE_*: how units of work (CTAs of whole ladders / chain groups)
                                      are placed on the SMs; results never depend on it                    */
} rwmpt_run_args_t;

int rwmpt_version(void);
const char* rwmpt_last_error(void);
uint64_t rwmpt_sizeof_run_args(void); /* ABI guard for FFI bindings */

/* RWM: n_temps must be 1.  PT: any n_temps >= 1, whole ladder resident in one CTA. */
int rwmpt_rwm_run(const rwmpt_run_args_t* args, void* cuda_stream);
int rwmpt_pt_run(const rwmpt_run_args_t* args, void* cuda_stream);

/* Number of swap sweeps a run with these step counts performs (host arithmetic only). */
int64_t rwmpt_count_swap_rounds(int64_t step_offset, int64_t n_steps, int64_t burn_in, int32_t swap_every);

/* Which (lanes_per_chain, elements_per_lane) the auto heuristic picks; returns lanes or <0. */
int rwmpt_pick_lanes(int32_t dim, int32_t n_temps, int64_t n_ladders, int32_t math_mode, int32_t* elems_per_lane);
/* Same for a full run description (target / proposal family and the lanes_per_chain request included: some BASELINE
 * workloads have tuned geometries); host arithmetic only. */
int rwmpt_pick_geometry(const rwmpt_run_args_t* args, int32_t* lanes_per_chain, int32_t* elems_per_lane);

/* Batched target log-density, replaces TorchTargetDistribution.log_density (interfaces/target_torch.py:32-43):
 * out[n] = log pi(x[n, :]). */
int rwmpt_log_density(const rwmpt_target_t* target, const float* x, int64_t n, float* out, int32_t math_mode,
                      void* cuda_stream);

/* Swap-probability estimate of the iterative temperature-ladder construction, replaces the inner estimate of
 * _construct_iterative_ladder (algorithms/pt_rwm_gpu_optimized.py:356-368) together with the targets' heuristic tempered
 * samplers draw_samples_torch (multimodal_torch.py:270-333, 532-565; rosenbrock_torch.py:224-248;
 * multivariate_normal_torch.py:101-121, 249-268):
 *   *sum_out += sum_{r < n} min(1, exp((beta_curr - beta_star) (log pi(x*_r) - log pi(x_r)))),
 * x_r ~ sampler(beta_curr), x*_r ~ sampler(beta_star), Philox subsequence row_id_base + r.  sum_out is a DEVICE double
 * the caller zeroes; the estimate is *sum_out / n.  RWMPT_ENOTSUP for families without such a sampler
 * (FullRosenbrock -- the reference raises there too --, HybridRosenbrock, NealFunnel, Hypercube, IIDGamma, IIDBeta). */
int rwmpt_swap_prob_estimate(const rwmpt_target_t* target, float beta_curr, float beta_star, int64_t n, uint64_t seed,
                             int64_t row_id_base, double* sum_out, void* cuda_stream);

/* Proposal increments, replaces ProposalDistribution.sample(n) (proposal_distributions/base.py:31-41):
 * out[n, dim]; row r uses Philox subsequence row_id_base + r. scale / dim_scale as in rwmpt_run_args_t
 * (scale is a single host float here). */
int rwmpt_proposal_sample(int32_t proposal_family, int32_t dim, float scale, const float* dim_scale, int64_t n,
                          uint64_t seed, int64_t row_id_base, float* out, void* cuda_stream);

/* Stand-alone adjacent-temperature swap sweep over HBM-resident ladders, replaces
 * _attempt_all_swaps (algorithms/pt_rwm_gpu_optimized.py:594-633).  One CTA keeps a whole ladder in shared
 * memory.  swap_uniforms [n_ladders, n_temps-1] nullable (then Philox(seed, ladder_id_base+l, round_index)). */
int rwmpt_pt_swap(float* state, float* logp, const float* beta, int64_t n_ladders, int32_t n_temps, int32_t dim,
                  int32_t swap_mode, const float* swap_uniforms, uint64_t seed, int64_t ladder_id_base,
                  int64_t round_index, unsigned char* swap_decisions, unsigned long long* swap_accepts,
                  void* cuda_stream);

/* ESJD / acceptance reduction over stored samples, replaces expected_squared_jump_distance_gpu
 * (algorithms/rwm_gpu_optimized.py:513-534, pt_rwm_gpu_optimized.py:772-789):
 * samples [n_chains, stride, dim]; for each chain esjd_out[c] = mean_{m=first+1..first+n-1} ||x_m - x_{m-1}||^2,
 * moved_out[c] (nullable) = number of those jumps that are non-zero. */
int rwmpt_esjd_reduce(const float* samples, int64_t n_chains, int64_t stride, int64_t first, int64_t n, int32_t dim,
                      double* esjd_out, unsigned long long* moved_out, void* cuda_stream);

/* Measures this GPU's FP32 FFMA and SFU (MUFU.EX2) issue peaks with dependent-free loops (synchronous; a few ms):
 * the denominators of the FP32 / SFU rooflines of SURVEY.md section 8(d).  fp32_tflops counts an FMA as 2 flops. */
int rwmpt_probe_peaks(double* fp32_tflops, double* sfu_gops);

/* Warp-instruction issue rate (warp-instr/s, whole GPU) of a dependent-free loop at a chosen occupancy:
 * kind 0 FFMA, 1 MUFU.EX2, 2 Philox rounds (IMAD/LOP3).  Diagnostic used by scripts/gpu_issue_probe.py. */
int rwmpt_probe_issue(int kind, int blocks, int threads, int iters, double* ops_per_s);

/* Philox4x32-10 known-answer hook (host in / host out, runs one device thread). */
int rwmpt_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/*
 * Host-buffer convenience entry: every pointer in `args` that is non-NULL is a HOST pointer; the call
 * allocates device buffers, copies in, runs rwmpt_pt_run, copies state / logp / accumulators / decisions and the
 * retained-sample rows THIS call wrote back (rows returned by earlier calls of a resumed run, and rows beyond
 * sample_rows, are left untouched on the host) and synchronises.  `device` is the CUDA ordinal; the caller's current
 * device is restored.  This is the end-to-end call the benchmark times (`e2e`).  h2d_bytes / d2h_bytes (nullable)
 * receive the bytes moved.
 */
int rwmpt_run_host(const rwmpt_run_args_t* args, int32_t device, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* RWMPT_H_ */
