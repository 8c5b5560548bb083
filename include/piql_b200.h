/*
 * piql_b200.h -- C ABI of the B200 (sm_100a) accelerator for PIQL.jl's data-parallel hot path.
 *
 * This is the drop-in boundary: a Julia host keeps PIQL's public API (generate_ν, generate_logz,
 * logz_updateV, generate_optimality_gap, plus a Monte-Carlo log Z estimator) and reaches these
 * entry points through `ccall`; see INTEGRATION.md for the Julia shim and julia/PIQLB200.jl.
 * Each entry point names the reference code it replaces (paths relative to the reference repo,
 * chelate/PIQL.jl).
 *
 * Conventions
 *   - Every function returns an int32 status, 0 = PIQL_OK.  On failure piql_last_error(ctx)
 *     returns a message owned by the library (valid until the next call on that context).
 *     No exceptions, no abort, nothing printed.
 *   - Per-cell arrays are column-major H x W like the Julia wall array (border included);
 *     cell (i, j), 0-based, has linear index i + H*j.  Wall cells are NaN in every output.
 *   - Per-(cell, action) arrays have the action fastest: Q[a + 4*(i + H*j)] (Julia Q[a,i,j]).
 *   - Host pointers are read/written only during the call; the library never keeps or frees
 *     them.  Calls block until the result is in the caller's buffers.  One caller at a time
 *     per context.
 *   - Actions 0..3 move by (+1,0) (-1,0) (0,+1) (0,-1) on (i,j)   src/Problems/GridWorld.jl:24-26
 *   - There is no CPU fallback: without a CUDA device piql_create fails.
 */
#ifndef PIQL_B200_H
#define PIQL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIQL_VERSION 100 /* 0.1.0 */

enum {
  PIQL_OK = 0,
  PIQL_ERR_ARG = 1,      /* bad argument (null pointer, size, range) */
  PIQL_ERR_STATE = 2,    /* call order: no grid loaded, no policy table, ... */
  PIQL_ERR_CUDA = 3,     /* CUDA runtime error; text in piql_last_error */
  PIQL_ERR_NO_DEVICE = 4,
  PIQL_ERR_UNSUPPORTED = 5,
  PIQL_ERR_NOT_CONVERGED = 6 /* max_sweeps reached; outputs hold the last iterate */
};

/* Arithmetic and storage type of the soft-VI sweeps.  PIQL_F32_FAST: fp32 with the hardware
 * ex2/lg2 approximations (MUFU) in the tiled sweep kernel; same storage as PIQL_F32. */
enum { PIQL_F64 = 0, PIQL_F32 = 1, PIQL_F32_FAST = 2 };

typedef struct piql_ctx piql_ctx;

/* keyword arguments of gridworld_ctrl   src/Problems/GridWorld.jl:90-95 */
typedef struct {
  double randomness; /* likelihood of a random action */
  double step_cost;  /* entropic cost of a step   (0.1 * reward_scale) */
  double reward;     /* reward of reaching the goal (1 * reward_scale) */
  double gamma;      /* discount */
} piql_grid_params;

/* Stop rule of the reference: sum_s |new(s) - old(s)| < tol   src/value_iteration.jl:74,82,112,149
 * (reference tol = 1e-12; sweeps are Jacobi here -- same fixed point). */
typedef struct {
  double tol;          /* <= 0: run exactly max_sweeps */
  int64_t max_sweeps;  /* hard cap */
  int32_t precision;   /* PIQL_F64 | PIQL_F32 | PIQL_F32_FAST */
  int32_t check_every; /* residual is evaluated every this many sweeps; <= 0: library default */
} piql_solve_opts;

typedef struct {
  int64_t sweeps;
  double residual;       /* sum |delta| of the last checked sweep */
  double device_seconds; /* CUDA-event time of the sweeps */
  int64_t backups;       /* state backups performed = sweeps * free cells */
  int32_t converged;
  int32_t kernel;        /* 1 = single-CTA resident solver, 2 = tiled multi-sweep kernel (4 sweeps per launch), 3 = batched resident, 4 = table, 5 = one launch per sweep (linear-space Z on large grids) */
} piql_solve_stats;

enum { PIQL_EST_SAMPLED_STOP = 0 /* estimator A */, PIQL_EST_DISCOUNTED = 1 /* estimator B */ };
enum { PIQL_START_ALL_FREE = 0 /* path p starts at free cell p mod nfree */, PIQL_START_LIST = 1 };

typedef struct {
  int64_t npaths;        /* paths in this call */
  int64_t path0;         /* global id of the first path (shards: rank r takes [path0, path0+npaths)) */
  uint64_t seed;         /* Philox4x32-10 key; counter = (step, 0, path id lo, path id hi) */
  int32_t estimator;     /* PIQL_EST_* */
  int32_t start_mode;    /* PIQL_START_* */
  int64_t horizon;       /* step cap per path, in [1, 2^30) (estimator B: the truncation horizon) */
  const int32_t* starts; /* PIQL_START_LIST: free (non-wall) cell indices; path p starts at starts[p mod nstarts] */
  int64_t nstarts;
  const float* injected; /* optional [npaths][injected_steps][4] uniforms replacing Philox */
  int64_t injected_steps;
} piql_mc_opts;

typedef struct {
  uint64_t transitions;  /* sampled (s,a,s') transitions */
  uint64_t paths;
  double device_seconds; /* rollout + reduction kernels, CUDA events */
  double rollout_seconds; /* rollout kernels only */
} piql_mc_stats;

/* ---- context -------------------------------------------------------------------------------
 * Replaces the reference's worker pool (addprocs/pmap/@spawn, scripts/distributed_experiments.jl
 * :3,26,28,43-46): one context drives one GPU; multi-GPU runs use one context per device (one
 * process or thread each) and combine statistics with the *_dev entry points below. */
int32_t piql_version(void);
int32_t piql_device_count(int32_t* n_out);
int32_t piql_create(piql_ctx** ctx_out, int32_t device_id);
int32_t piql_destroy(piql_ctx* ctx);
const char* piql_last_error(const piql_ctx* ctx);
/* cudaStream_t all of this context's work is launched on (for event timing / interop) */
int32_t piql_stream(piql_ctx* ctx, void** stream_out);

/* ---- problem --------------------------------------------------------------------------------
 * Tabulated gridworld_ctrl(gw; randomness, step_cost, reward, γ)   src/Problems/GridWorld.jl:90-121
 * walls: the Gridworld.walls Bool array (:35-38), 1 = wall, border included; goal 0-based. */
int32_t piql_grid_load(piql_ctx* ctx, const uint8_t* walls_colmajor, int64_t H, int64_t W,
                       int64_t goal_i, int64_t goal_j, const piql_grid_params* params);
int32_t piql_grid_free_count(piql_ctx* ctx, int64_t* nfree_out);

/* ---- soft value iteration -------------------------------------------------------------------
 * generate_ν(ctrl)   src/value_iteration.jl:62-87, truevalue_recursion src/PIQL.jl:95-101.
 * reward_scale multiplies the reward function (modify(ctrl, reward_function = c*r),
 * scripts/experiments.jl:26, test/runtests.jl:7-8). */
int32_t piql_solve_nu(piql_ctx* ctx, double reward_scale, const piql_solve_opts* opts, double* nu_out,
                      piql_solve_stats* stats_out);
/* generate_logz(ctrl, V; β)   src/value_iteration.jl:89-117, logz_recursion src/PIQL.jl:125-130.
 * Q: optional per-(cell,action) table; NULL = the V-determined Q, (s,a)->controlV(ctrl,V,s,a). */
int32_t piql_solve_logz(piql_ctx* ctx, const double* V, const double* Q, double beta,
                        const piql_solve_opts* opts, double* logz_out, piql_solve_stats* stats_out);
/* logz_updateV(ctrl, V; β) = V + logz/β   src/value_iteration.jl:119-124.  logz_out optional. */
int32_t piql_update_v(piql_ctx* ctx, const double* V, double beta, const piql_solve_opts* opts,
                      double* vnew_out, double* logz_out, piql_solve_stats* stats_out);
/* generate_z(ctrl, V; β) and z_updateV(ctrl, V; β) = V + log(Z)/β   src/value_iteration.jl:23-59 over z_recursion
 * src/PIQL.jl:109-115: the linear-space twin of the log Z solve (Z starts at 1, linear damping η = ½, stop rule on
 * Σ|log new − log old|; the reference calls it "not numerically stable enough" and so it serves as a cross-check:
 * log(Z) agrees with piql_solve_logz wherever exp does not under/overflow).  Swept by the per-cell kernels
 * (stats.kernel 1 or 5).  Q and z_out optional as above. */
int32_t piql_solve_z(piql_ctx* ctx, const double* V, const double* Q, double beta,
                     const piql_solve_opts* opts, double* z_out, piql_solve_stats* stats_out);
int32_t piql_z_update_v(piql_ctx* ctx, const double* V, double beta, const piql_solve_opts* opts,
                        double* vnew_out, double* z_out, piql_solve_stats* stats_out);
/* generate_optimality_gap(ctrl, V; ν)   src/value_iteration.jl:129-169 */
int32_t piql_solve_gap(piql_ctx* ctx, const double* V, const double* Q, const double* nu,
                       const piql_solve_opts* opts, double* gap_out, piql_solve_stats* stats_out);
/* The experiment protocol in two calls (run_experiment, scripts/experiments.jl:20-39; its distributed twin
 * scripts/distributed_experiments.jl:15-33 fans the same work out with pmap over beta and over iterates):
 *   piql_learning_traj: for every beta, learning_traj(ctrl, V0; beta, n = nupdates) (scripts/experiments.jl:10-18):
 *     traj_out[b][0] = V0, traj_out[b][k] = logz_updateV(ctrl, traj_out[b][k-1]; beta_b);  nbeta*(nupdates+1)*H*W doubles,
 *     stats_out (optional) nbeta*nupdates entries.  The chains are independent: one CTA each, in one launch.
 *   piql_solve_gap_batch: generate_optimality_gap(ctrl, Vs[i]; nu) for m value functions at once (m*H*W doubles in and out).
 * Problems too large for one CTA's shared memory run the same solves one after the other with the tiled kernel. */
int32_t piql_learning_traj(piql_ctx* ctx, const double* V0, const double* betas, int32_t nbeta, int32_t nupdates,
                           const piql_solve_opts* opts, double* traj_out, piql_solve_stats* stats_out);
int32_t piql_solve_gap_batch(piql_ctx* ctx, const double* Vs, int32_t m, const double* nu, const piql_solve_opts* opts,
                             double* gaps_out, piql_solve_stats* stats_out);
/* ---- tabulated ControlProblems ---------------------------------------------------------------
 * Any finite problem behind the closure interface of src/PIQL.jl:12-23 (e.g. src/Problems/BinaryBandit.jl), tabulated by
 * the host: states 0..nstates-1, actions 0..nactions-1 (<= 16), and for every (state, action) row s*nactions + a the
 * entries row_ptr[row] .. row_ptr[row+1]-1 of (successor, probability, reward_function(s, a, s')):
 *   propagator_average(s,a,f) = sum_e prob[e] f(succ[e])      logpropagatorexp(s,a,f) = log sum_e prob[e] exp f(succ[e])
 * log_prior[row] = log action_prior(s, a); terminal[s] = terminal_condition(s); gamma in (0, 1] (1: finite-horizon problems).
 * Solvers as for the grid (V-determined Q, Float64, Jacobi, stop rule of the reference); arrays are nstates long. */
typedef struct {
  int64_t nstates;
  int32_t nactions;
  const int64_t* row_ptr;
  const int32_t* succ;
  const double* prob;
  const double* reward;
  const double* log_prior;
  const uint8_t* terminal;
  double gamma;
} piql_table_problem;
int32_t piql_table_load(piql_ctx* ctx, const piql_table_problem* problem);
/* A tabulated Q for the loaded table problem: Q[s * nactions + a] (any Q callable of src/PIQL.jl:72-89 evaluated on the
 * tabulated states), used by the log Z / gap solves, piql_table_update_v and piql_table_mc_logz that follow -- Qnormalization,
 * controlQ, action_probability and fitness are taken from it, controlV still from V.  NULL returns to the V-determined Q
 * (s, a) -> controlV(ctrl, V, s, a) of value_iteration.jl:97,157.  Copied: the caller keeps ownership.  piql_table_load
 * clears it. */
int32_t piql_table_set_q(piql_ctx* ctx, const double* Q);
/* kind 0: generate_ν (reward_function scaled by reward_scale); 1: generate_logz(ctrl, V; β); 2: generate_optimality_gap(ctrl, V; ν) */
int32_t piql_table_solve(piql_ctx* ctx, int32_t kind, const double* V, const double* nu, double beta, double reward_scale,
                         const piql_solve_opts* opts, double* x_out, piql_solve_stats* stats_out);
/* logz_updateV(ctrl, V; β) on the tabulated problem; logz_out optional */
int32_t piql_table_update_v(piql_ctx* ctx, const double* V, double beta, const piql_solve_opts* opts, double* vnew_out,
                            double* logz_out, piql_solve_stats* stats_out);

/* Monte-Carlo estimate of log Z on the tabulated problem (sampled-stop estimator; opts->starts are state indices, or
 * PIQL_START_ALL_FREE = path p starts at state p mod nstates).  Per step of path p one Philox4x32-10 call, counter
 * (step, 0, p_lo, p_hi): word 0 -> action (StatsBase walk over the running action probabilities), word 1 -> stop test
 * u >= gamma, word 2 -> successor (the same walk over the running transition probabilities of the row).
 * logz_sa_out / count_sa_out: nstates*nactions, logz_s_out: nstates (NaN / 0 where no path started); any may be NULL.
 * cdf_out, bf_out (nstates*nactions floats) and cum_out (nnz floats) optionally export the sampler's tables (running action
 * probabilities, beta*fitness, running transition probabilities) for trajectory-parity checks; the per-path records are
 * read with piql_mc_records (meta = first action | length << 4). */
int32_t piql_table_mc_logz(piql_ctx* ctx, const double* V, double beta, const piql_mc_opts* opts, double* logz_sa_out,
                           uint64_t* count_sa_out, double* logz_s_out, piql_mc_stats* stats_out, float* cdf_out,
                           float* bf_out, float* cum_out);

/* One Jacobi sweep from x_in (kind: 0 nu, 1 logz, 2 gap, 3 z), for parity tests of the backup itself. */
int32_t piql_sweep_once(piql_ctx* ctx, int32_t kind, const double* V, const double* Q, const double* nu,
                        double beta, double reward_scale, int32_t precision, const double* x_in,
                        double* x_out, double* residual_out);

/* The 16-byte form of the current policy records, which the production rollout kernel reads on grids whose 32-byte table
 * does not stay in L2 (>= 2^22 cells; V-determined Q only, where beta * fitness does not depend on the action):
 * per cell four 32-bit words {c0 | (mask & 3) << 30, c1 | (mask >> 2) << 30, c2, bits of beta * fitness(s, action 0)} --
 * exactly the 32-byte record's c0, c1, c2, mask and first fitness entry.  *in_use_out = 1 when the rollouts use it (then
 * table16_out, if non-null, receives H*W*4 words), 0 when they read the 32-byte records. */
int32_t piql_policy_table16(piql_ctx* ctx, uint32_t* table16_out, int32_t* in_use_out);
/* piql_policy_table without blocking and without serialising behind work already enqueued: V (and Q) are uploaded and the
 * table is built on the library's upload stream into a spare buffer while rollouts enqueued earlier still read the current
 * table; calls enqueued after this one use the new table (the context's stream waits for it on the device).  For callers
 * that iterate: the upload of iteration k + 1 runs under the rollout of iteration k.  V / Q should be pinned host memory and
 * must stay unchanged until a later blocking call on this context returns (e.g. piql_mc_finalize_wait of that iteration). */
int32_t piql_policy_table_async(piql_ctx* ctx, const double* V, const double* Q, double beta);
/* ---- Monte-Carlo path sampling --------------------------------------------------------------
 * Replaces choose_action / path_transitions / fitness_reward_estimate   src/path_functions.jl:12-41
 * and the propagator closure src/Problems/GridWorld.jl:113-115.
 *
 * piql_policy_table builds, on the device, the per-cell record the sampler reads (32 bytes):
 *   float c[3]   running sums of action_probability(ctrl,Q,s,a)/sum  (StatsBase walk cw < u)
 *   uint32 mask  bit a set <=> take_step(s,a) is blocked by a wall
 *   float bf[4]  beta * fitness(ctrl,V,Q,s,a)                         src/PIQL.jl:87-89
 * table_out (optional) receives H*W*8 floats.  piql_policy_table_set uploads a caller's table. */
int32_t piql_policy_table(piql_ctx* ctx, const double* V, const double* Q, double beta, float* table_out);
int32_t piql_policy_table_set(piql_ctx* ctx, const float* table, double v_goal);

/* Monte-Carlo estimate of log Z (the fixed point of logz_recursion, src/PIQL.jl:125-130) from
 * npaths rollouts under pi_Q, bucketed by (start cell, first action):
 *   logz_sa_out[a + 4*cell]  log mean path weight of the bucket (NaN when empty)   H*W*4
 *   count_sa_out[a + 4*cell] paths in the bucket                                   H*W*4
 *   logz_s_out[cell]         log mean path weight over the cell's four buckets     H*W
 * Any output pointer may be NULL.  Builds the policy table from V (and Q) first. */
int32_t piql_mc_logz(piql_ctx* ctx, const double* V, const double* Q, double beta, const piql_mc_opts* opts,
                     double* logz_sa_out, uint64_t* count_sa_out, double* logz_s_out,
                     piql_mc_stats* stats_out);

/* The Monte-Carlo form of logz_updateV (src/value_iteration.jl:119-124), niter times, device-resident: each iteration
 * rebuilds the policy table from the current V, rolls out opts->npaths paths (seed opts->seed + iteration), reduces them and
 * applies V(s) += log Z^(s) / beta at every start cell that received paths.  vnew_out: H*W doubles, NaN at walls. */
int32_t piql_mc_update_v(piql_ctx* ctx, const double* V, double beta, const piql_mc_opts* opts, int32_t niter,
                         double* vnew_out, piql_mc_stats* stats_out);

/* fitness_reward_estimate / direct_reward_estimate   src/path_functions.jl:34-50 over path_transitions :23-32, for
 * opts->npaths paths at once: exactly nsteps sampled transitions per path under π_Q (the reference's walk has no stop
 * rule and does not treat the goal specially; estimator / horizon in opts are ignored), and after every transition
 *   fitness_out[p][t] = V(s0) + Σ_{τ≤t} fitness(s_τ, a_τ)
 *   direct_out[p][t]  = V(s0) + Σ_{τ≤t} [reward_function(s_τ, a_τ, s'_τ) − kl_cost(s_τ)]      (sampled reward)
 * in fp32 ([npaths][nsteps] each).  The reference file does not load as written (`ctrl.reward`, a missing `return out`);
 * these are the evident intent, as the oracle restates them.  Uniforms: the estimator's contract (word 1 of a step, the
 * stop test, is drawn and unused).  Optional outputs: triples_out [npaths][nsteps][3] (s, a, s'); table_out (32 B per
 * cell, the β = 1 sampler records) and kl_out (one float per cell) -- what the kernel read, for bit-exact checks.
 * Replaces the current policy table: call piql_policy_table again before piql_mc_rollout_dev. */
int32_t piql_path_estimates(piql_ctx* ctx, const double* V, const double* Q, const piql_mc_opts* opts, int32_t nsteps,
                            float* fitness_out, float* direct_out, int32_t* triples_out, float* table_out, float* kl_out);

/* Debug export for bit-exact trajectory parity: the (s, a, s') triples of each path (int32 cell
 * indices, -1 padded, [npaths][max_steps][3]) and the per-path record: logw (float) and
 * meta = a0 | (len << 2).  Uses the current policy table. */
int32_t piql_trace_paths(piql_ctx* ctx, const piql_mc_opts* opts, int64_t max_steps, int32_t* triples_out,
                         float* logw_out, uint32_t* meta_out);

/* ---- device-resident entry points (bench + multi-GPU) ---------------------------------------
 * The per-(start slot, action) accumulators (lse_max float, lse_sum float, count uint32, each
 * nslots*4 long, slot = index into the start list / free-cell list) stay on the device so that
 * shards can be combined by an all-reduce before they are finalised:
 *     piql_mc_reset_dev -> piql_mc_rollout_dev (xN) -> [allreduce MAX on lse_max into gmax ->
 *     piql_mc_rescale_dev(gmax) -> allreduce SUM on lse_sum and count] -> piql_mc_finalize
 * All of them are asynchronous on the context's stream except finalize. */
int32_t piql_mc_reset_dev(piql_ctx* ctx, const piql_mc_opts* opts);
int32_t piql_mc_rollout_dev(piql_ctx* ctx, const piql_mc_opts* opts);
int32_t piql_mc_stats_dev(piql_ctx* ctx, void** lse_max_dev, void** lse_sum_dev, void** count_dev,
                          int64_t* n_out);
int32_t piql_mc_rescale_dev(piql_ctx* ctx, const void* global_max_dev);
int32_t piql_mc_finalize(piql_ctx* ctx, double* logz_sa_out, uint64_t* count_sa_out, double* logz_s_out,
                         piql_mc_stats* stats_out);
/* Asynchronous finalisation, for iterating callers: the accumulators are snapshot on the context's stream (a device copy),
 * then finalised and copied to the caller's host buffers on a second stream while the context's stream goes on with the
 * next iteration (piql_policy_table / piql_mc_reset_dev / piql_mc_rollout_dev ...).  The host buffers (pinned memory makes
 * the copy truly asynchronous) must stay valid and unread until piql_mc_finalize_wait returns; at most one finalisation is
 * in flight (a second call waits for the first on the device).  f32_outputs = 0: double / uint64 outputs as piql_mc_finalize;
 * 1: float / uint32 outputs of the same shapes (36 instead of 72 bytes per cell over PCIe).  Any output may be null.
 * Replaces the master collecting the workers' results (scripts/distributed_experiments.jl:49-58). */
int32_t piql_mc_finalize_async(piql_ctx* ctx, void* logz_sa_out, void* count_sa_out, void* logz_s_out, int32_t f32_outputs);
int32_t piql_mc_finalize_wait(piql_ctx* ctx, piql_mc_stats* stats_out);
/* The exchange step as ONE kernel over NVLink peer memory (one process per GPU, all on one box): rank r pulls
 * slice r of every rank's three accumulator arrays through CUDA-IPC peer mappings, combines the triples and
 * pushes the result into every rank's arrays (replaces the three all-reduces + rescale above).
 *   piql_peer_export (after piql_mc_reset_dev) -> exchange the handle blobs between the ranks (any transport)
 *   -> piql_peer_import(rank, world, all blobs in rank order) -> piql_mc_allreduce_peer once per iteration.
 * Re-export after a reset that changed the start list size.  Every rank must make the same calls. */
typedef struct {
  uint8_t lse_max[64], lse_sum[64], count[64], flags[64]; /* cudaIpcMemHandle_t */
  int64_t n;                                              /* accumulator length (slots * 4) */
} piql_peer_handles;
int32_t piql_peer_export(piql_ctx* ctx, piql_peer_handles* out);
int32_t piql_peer_import(piql_ctx* ctx, int32_t rank, int32_t world, const piql_peer_handles* all);
int32_t piql_mc_allreduce_peer(piql_ctx* ctx);
int32_t piql_peer_close(piql_ctx* ctx);

/* CUDA-event durations of the rollout and reduction kernels launched since the previous call of
 * this function (or since context creation), summed over launches.  Synchronises the stream. */
int32_t piql_mc_timing(piql_ctx* ctx, double* rollout_seconds, double* reduce_seconds, int64_t* launches);
/* Sampled transitions of every rollout on this context since creation (synchronises the stream). */
int32_t piql_mc_transitions_total(piql_ctx* ctx, uint64_t* total_out);
/* Per-path records of the last piql_mc_rollout_dev call, copied to the host (tests). */
int32_t piql_mc_records(piql_ctx* ctx, float* logw_out, uint32_t* meta_out, int64_t n);

/* Device-resident soft-VI for timing: runs `sweeps` Jacobi sweeps of kind (0 nu, 1 logz, 2 gap)
 * on the context's stream from the current device state without touching the host.
 * piql_vi_prepare_dev uploads V/Q/nu (as needed) and resets the iterate. */
int32_t piql_vi_prepare_dev(piql_ctx* ctx, int32_t kind, const double* V, const double* Q, const double* nu,
                            double beta, double reward_scale, int32_t precision);
/* Shape of the tiled soft-VI kernel for the loaded grid (grids above 8192 cells): plan_out[8] =
 * {interior columns per tile, warps per CTA, tile rows, tile columns, CTAs per SM, shared bytes per CTA,
 *  sweeps per round (4 or 8: the temporal-blocking depth the library timed fastest), flags: bit 0 = all tiles are
 *  resident at once (persistent kernel: one cooperative launch per batch), bits 1..3 = constant planes the linear fast fp32
 *  sweeps keep on chip (log Z: 4, gap: 5; 0 = read from global memory), bit 4 = the column-resident schedule (grids of at
 *  most two 128-row strips per warp: one CTA per SM keeps a slab of whole columns in shared memory for the whole batch of
 *  sweeps -- a cooperative launch per batch -- and the fields read {columns per slab, warps per CTA, 1, slabs, 1, shared
 *  bytes, -, flags})}. */
int32_t piql_vi_tile_plan(piql_ctx* ctx, int32_t kind, int32_t precision, int32_t* plan_out);
/* Replaces the current iterate (warm start / resume from a caller-kept array; wall cells ignored). */
int32_t piql_vi_set_iterate(piql_ctx* ctx, const double* x);
int32_t piql_vi_sweeps_dev(piql_ctx* ctx, int64_t sweeps);
int32_t piql_vi_fetch(piql_ctx* ctx, double* x_out, double* residual_out);

/* Multi-GPU soft VI: the grid is cut into slabs of tile columns, one per rank (one process per GPU on one box); after every
 * round of sweeps a rank pushes its edge columns into its neighbours' arrays over NVLink peer memory.  Jacobi sweeps and
 * tiling-independent values make the result bit-identical to the single-GPU solve for any number of ranks.
 *   piql_vi_prepare_dev [-> piql_vi_set_iterate] -> piql_vi_shard_begin(rank, world, &mine) -> exchange the handle blobs
 *   -> piql_vi_shard_connect(all blobs in rank order) -> barrier -> piql_vi_sweeps_dev (xN; the residual it leaves is this
 *   rank's share: sum over ranks) -> piql_vi_fetch (only the columns [col0, col1) of the blob are this rank's) -> piql_vi_shard_end */
typedef struct {
  uint8_t x0[64], x1[64], flags[64]; /* cudaIpcMemHandle_t of the padded ping-pong arrays and the halo counters */
  int64_t np;                        /* padded array length */
  int32_t tj, tiles_j;               /* tile shape */
  int64_t col0, col1;                /* columns of the H x W array this rank owns */
} piql_vi_shard_handles;
int32_t piql_vi_shard_begin(piql_ctx* ctx, int32_t rank, int32_t world, piql_vi_shard_handles* out);
int32_t piql_vi_shard_connect(piql_ctx* ctx, const piql_vi_shard_handles* all);
int32_t piql_vi_shard_end(piql_ctx* ctx);
/* Diagnostics of the halo exchange since the last call (or piql_vi_shard_begin): microseconds this rank's edge CTAs spent
 * waiting for a neighbour's counter, summed over the waiting CTAs, and the number of such waits (edge CTAs x rounds). */
int32_t piql_vi_shard_wait_stats(piql_ctx* ctx, double* wait_us_out, int64_t* waits_out);

/* Kernels launched by this context since creation (for the bench's gpu_launches count). */
int32_t piql_launch_count(piql_ctx* ctx, int64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* PIQL_B200_H */
