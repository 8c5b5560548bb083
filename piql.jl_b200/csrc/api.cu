// api.cu -- the C ABI (include/piql_b200.h): context, device buffers, call sequencing.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/piql_b200.h"
#include "internal.h"

using namespace piql;

struct piql_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  long long launches = 0;

  // problem
  bool has_grid = false;
  GridDev g{};
  long long nfree = 0;
  uint8_t* d_walls = nullptr;
  uint8_t* d_cell = nullptr;
  std::vector<uint8_t> h_walls;  // host copy of the wall bytes (NaN pattern of host-composed outputs)
  bool h_walls_nan(long long s) const { return h_walls[(size_t)s] != 0; }
  int* d_free = nullptr;  // free cells in index order

  // soft VI
  long long vi_cap = 0;  // cells the buffers below were sized for
  ViState vs{};
  bool vs_ready = false;
  double *d_in0 = nullptr, *d_in1 = nullptr, *d_in4 = nullptr, *d_out = nullptr;  // double staging
  void* d_Q = nullptr;
  void* d_nu = nullptr;
  double* d_scalars = nullptr;  // [0] residual, [1..2] resident solver output
  double* h_scalars = nullptr;  // pinned mirror
  long long vi_sweeps_done = 0;
  TileBuffers tb;  // padded copy of the iterate for the tiled multi-sweep kernel
  PeerState peer;  // NVLink peer mappings of the other ranks' accumulators
  ViPeer vpeer;    // ... and of their padded soft-VI arrays (column-slab sharding)

  // tabulated (non-grid) problem
  bool has_table_problem = false;
  TableDev tab{};
  long long tab_nnz = 0;
  std::vector<void*> tab_bufs;  // every device allocation of the table path
  double* tb_Q = nullptr;        // optional tabulated Q [ns * na] (piql_table_set_q); nullptr: the V-determined Q
  bool tb_has_q = false;
  double *tb_V = nullptr, *tb_nu = nullptr, *tb_x[2] = {nullptr, nullptr}, *tb_aux = nullptr, *tb_aux1 = nullptr,
         *tb_partial = nullptr, *tb_out = nullptr;

  // Monte-Carlo
  float4* d_table = nullptr;
  uint4* d_table16 = nullptr;        // the same policy as 16-byte records (large grids, V-determined Q), read by the production rollout
  uint4* d_table16_spare = nullptr;
  bool t16_valid = false, t16_valid_spare = false;
  bool has_table = false;
  double v_goal = 0.0;
  uint2* d_rec = nullptr;        // per-path records of the last rollout launch {log-weight bits, a0 | len << 2}
  long long rec_cap = 0;
  float* d_logw = nullptr;       // split copies: the table path's records; de-interleaved d_rec for piql_mc_records
  uint32_t* d_meta = nullptr;
  long long split_cap = 0;
  bool records_split = false;    // the last rollout was the table sampler's: d_logw / d_meta already hold its records
  int* d_starts = nullptr;       // current slot list (points at d_free or d_starts_own)
  int* d_starts_own = nullptr;
  long long starts_own_cap = 0;
  long long nstarts = 0;
  float *acc_max = nullptr, *acc_sum = nullptr;
  uint32_t* acc_cnt = nullptr;
  long long acc_cap = 0;
  float* d_scratch = nullptr;
  size_t scratch_cap = 0;
  unsigned long long* d_transitions = nullptr;
  unsigned long long* h_transitions = nullptr;
  float* d_inj = nullptr;
  size_t inj_cap = 0;
  unsigned long long paths_done = 0;
  // asynchronous policy table (piql_policy_table_async): the next table is uploaded and built on an upload stream into a
  // spare buffer while rollouts enqueued earlier still read the current one
  cudaStream_t ustream = nullptr;
  float4* d_table_spare = nullptr;
  double *d_upV = nullptr, *d_upQ = nullptr;   // upload staging of V (n) and Q (4n)
  long long up_cap = 0, upq_cap = 0;
  cudaEvent_t ev_retire_cur = nullptr;    // recorded on the stream when the CURRENT table's buffer was last made current ... see below
  cudaEvent_t ev_retire_spare = nullptr;  // everything that may read the SPARE buffer had been enqueued before this event
  cudaEvent_t ev_table_ready = nullptr;
  bool spare_retired = false;
  // asynchronous finalisation (piql_mc_finalize_async): a snapshot of the accumulators is finalised and copied to the
  // host on a second stream while the context's stream goes on with the next iteration
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_snap = nullptr, ev_done = nullptr;
  float *snap_max = nullptr, *snap_sum = nullptr;
  uint32_t* snap_cnt = nullptr;
  long long snap_cap = 0;
  unsigned char* fin_out = nullptr;   // device staging of the outputs
  size_t fin_cap = 0;
  unsigned long long* h_fin_transitions = nullptr;  // pinned
  unsigned long long fin_paths = 0;
  bool fin_pending = false;
  // per-launch timing: triples (before rollout, after rollout, after reduce) since the last reset
  std::vector<cudaEvent_t> tev;
  size_t tev_used = 0;
};

namespace {

constexpr long long MAX_RECORDS_PER_LAUNCH = 1LL << 28;  // 2 GiB of per-path records
constexpr size_t MAX_TIMING_EVENTS = 3 * 4096;

struct LaunchScope {
  piql_ctx* c;
  long long l0;
  explicit LaunchScope(piql_ctx* ctx) : c(ctx), l0(g_launches) {}
  ~LaunchScope() { c->launches += g_launches - l0; }
};

int fail(piql_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(ctx, PIQL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));      \
  } while (0)
#define CK_LAST() CK(cudaGetLastError())
#define NEED(cond, code, msg) \
  do { if (!(cond)) return fail(ctx, code, msg); } while (0)

template <typename T> cudaError_t grow(T*& p, long long& cap, long long want) {
  if (want <= cap) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  cudaError_t e = cudaMalloc((void**)&p, sizeof(T) * (size_t)want);
  if (e == cudaSuccess) cap = want;
  return e;
}
template <typename T> cudaError_t grow(T*& p, size_t& cap, size_t want) {
  long long c = (long long)cap;
  cudaError_t e = grow(p, c, (long long)want);
  cap = (size_t)c;
  return e;
}

void peer_close(piql_ctx* ctx) {
  PeerState& ps = ctx->peer;
  for (int i = 0; i < ps.nopened; i++) cudaIpcCloseMemHandle(ps.opened[i]);
  ps.nopened = 0;
  ps.valid = false;
}

void vpeer_close(piql_ctx* c) {
  ViPeer& vp = c->vpeer;
  for (int i = 0; i < vp.nopened; i++) cudaIpcCloseMemHandle(vp.opened[i]);
  vp.nopened = 0;
  vp.valid = false;
  c->tb.sharded = false;
}

void free_table(piql_ctx* c) {
  for (void* p : c->tab_bufs) cudaFree(p);
  c->tab_bufs.clear();
  c->has_table_problem = false;
}

void free_vi(piql_ctx* c) {
  for (void* p : {c->vs.x[0], c->vs.x[1], c->vs.V, c->vs.aux4, c->vs.aux1, (void*)c->vs.partial, (void*)c->d_in0,
                  (void*)c->d_in1, (void*)c->d_in4, (void*)c->d_out, c->d_Q, c->d_nu})
    if (p) cudaFree(p);
  c->vs = ViState{};
  c->d_in0 = c->d_in1 = c->d_in4 = c->d_out = nullptr;
  c->d_Q = c->d_nu = nullptr;
  c->vi_cap = 0;
  c->vs_ready = false;
  tile_free(c->tb);
}

int ensure_vi(piql_ctx* ctx) {
  const long long n = ctx->g.n;
  if (ctx->vi_cap >= n) return PIQL_OK;
  free_vi(ctx);
  const size_t e8 = 8 * (size_t)n;  // sized for double; float runs use half
  CK(cudaMalloc(&ctx->vs.x[0], e8));
  CK(cudaMalloc(&ctx->vs.x[1], e8));
  CK(cudaMalloc(&ctx->vs.V, e8));
  CK(cudaMalloc(&ctx->vs.aux4, 4 * e8));
  CK(cudaMalloc(&ctx->vs.aux1, e8));
  CK(cudaMalloc(&ctx->d_Q, 4 * e8));
  CK(cudaMalloc(&ctx->d_nu, e8));
  CK(cudaMalloc((void**)&ctx->d_in0, e8));
  CK(cudaMalloc((void**)&ctx->d_in1, e8));
  CK(cudaMalloc((void**)&ctx->d_in4, 4 * e8));
  CK(cudaMalloc((void**)&ctx->d_out, e8));
  const int nb = (int)((n + SWEEP_THREADS - 1) / SWEEP_THREADS);
  CK(cudaMalloc((void**)&ctx->vs.partial, sizeof(double) * nb));
  ctx->vs.nblocks = nb;
  ctx->vs.residual = ctx->d_scalars;
  ctx->vi_cap = n;
  return PIQL_OK;
}

// Upload V/Q/nu, convert to the sweep precision, reset the iterate, precompute per-cell constants.
int vi_prepare(piql_ctx* ctx, int kind, const double* V, const double* Q, const double* nu, double beta,
               double reward_scale, int precision) {
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  NEED(kind >= 0 && kind <= 3, PIQL_ERR_ARG, "kind must be 0 (nu), 1 (logz), 2 (gap) or 3 (z)");
  NEED(precision == PIQL_F64 || precision == PIQL_F32 || precision == PIQL_F32_FAST, PIQL_ERR_ARG,
       "precision must be PIQL_F64, PIQL_F32 or PIQL_F32_FAST");
  const bool fast = precision == PIQL_F32_FAST;
  if (fast) precision = PIQL_F32;
  NEED(kind == KIND_NU || V != nullptr, PIQL_ERR_ARG, "V is required for logz / gap");
  NEED(kind != KIND_GAP || nu != nullptr, PIQL_ERR_ARG, "nu is required for the optimality gap");
  NEED(kind == KIND_NU || (beta == beta), PIQL_ERR_ARG, "beta is NaN");
  int rc = ensure_vi(ctx);
  if (rc) return rc;
  const int n = ctx->g.n;
  cudaStream_t st = ctx->stream;
  ViState& vs = ctx->vs;
  vs.precision = precision; vs.kind = kind; vs.beta = beta; vs.reward_scale = reward_scale; vs.cur = 0;
  vs.fast = fast;
  ctx->tb.active = false;
  if (ctx->tb.sharded) vpeer_close(ctx);  // a new solve starts unsharded; piql_vi_shard_begin shards it again
  vs.nblocks = (n + SWEEP_THREADS - 1) / SWEEP_THREADS;
  vs.v_goal = 0.0;
  vs.residual = ctx->d_scalars;
  if (V) {
    const double vg = V[ctx->g.goal];
    NEED(vg == vg, PIQL_ERR_ARG, "V(goal) is NaN");
    vs.v_goal = vg;
    CK(cudaMemcpyAsync(ctx->d_in0, V, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    launch_convert_in(ctx->d_in0, vs.V, n, precision, st);
  }
  const void* dQ = nullptr;
  if (Q && kind != KIND_NU) {
    CK(cudaMemcpyAsync(ctx->d_in4, Q, 32 * (size_t)n, cudaMemcpyHostToDevice, st));
    launch_convert_in(ctx->d_in4, ctx->d_Q, 4 * n, precision, st);
    dQ = ctx->d_Q;
  }
  if (nu && kind == KIND_GAP) {
    CK(cudaMemcpyAsync(ctx->d_in1, nu, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    launch_convert_in(ctx->d_in1, ctx->d_nu, n, precision, st);
  }
  const double x_default = kind == KIND_Z ? 1.0 : 0.0;  // Growable default x -> 0.0 (x -> 1.0 for Z)  value_iteration.jl:44,79,109,146
  launch_fill(vs.x[0], x_default, n, precision, st);
  launch_fill(vs.x[1], x_default, n, precision, st);
  // per-cell constants from the Float64 uploads (d_in0 = V, d_in4 = Q, d_in1 = nu), rounded once to the sweep precision
  launch_prep(ctx->g, ctx->d_cell, vs, ctx->d_in0, dQ ? ctx->d_in4 : nullptr, ctx->d_in1, st);
  CK_LAST();
  ctx->vs_ready = true;
  ctx->vi_sweeps_done = 0;
  return PIQL_OK;
}

int vi_solve(piql_ctx* ctx, const piql_solve_opts* opts, double* out, piql_solve_stats* stats) {
  NEED(opts != nullptr, PIQL_ERR_ARG, "opts is null");
  NEED(opts->max_sweeps > 0, PIQL_ERR_ARG, "max_sweeps must be positive");
  const int n = ctx->g.n;
  cudaStream_t st = ctx->stream;
  ViState& vs = ctx->vs;
  const double tol = opts->tol;
  long long sweeps = 0;
  double residual = INFINITY;
  int kernel = 0;
  CK(cudaEventRecord(ctx->ev0, st));
  if (n <= RESIDENT_MAX_CELLS) {
    kernel = 1;
    launch_resident_solve(ctx->g, ctx->d_cell, vs, tol > 0 ? tol : -1.0, opts->max_sweeps, ctx->d_scalars + 1, st);
    CK_LAST();
    CK(cudaMemcpyAsync(ctx->h_scalars + 1, ctx->d_scalars + 1, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    sweeps = (long long)ctx->h_scalars[1];
    residual = ctx->h_scalars[2];
  } else if (vs.kind == KIND_Z) {
    // the linear-space cross-check has no tiled kernel: one launch per sweep, residual read back every check_every sweeps
    kernel = 5;
    const long long every = opts->check_every > 0 ? opts->check_every : 64;
    while (sweeps < opts->max_sweeps) {
      const long long batch = opts->max_sweeps - sweeps < every ? opts->max_sweeps - sweeps : every;
      for (long long i = 0; i < batch; i++) launch_sweep(ctx->g, ctx->d_cell, vs, i == batch - 1, st);
      CK_LAST();
      sweeps += batch;
      CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      residual = ctx->h_scalars[0];
      if (!(residual == residual)) return fail(ctx, PIQL_ERR_CUDA, "soft-VI residual is NaN (diverged)");
      if (tol > 0 && residual < tol) break;
    }
  } else {
    kernel = 2;
    CK(tile_begin(ctx->g, ctx->d_cell, vs, ctx->tb, ctx->num_sms, st));
    const long long every = opts->check_every > 0 ? opts->check_every : 64;
    while (sweeps < opts->max_sweeps) {
      long long batch = opts->max_sweeps - sweeps < every ? opts->max_sweeps - sweeps : every;
      CK(tile_sweeps(ctx->g, vs, ctx->tb, batch, true, st));
      sweeps += batch;
      CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      residual = ctx->h_scalars[0];
      if (!(residual == residual)) return fail(ctx, PIQL_ERR_CUDA, "soft-VI residual is NaN (diverged)");
      if (tol > 0 && residual < tol) break;
    }
    tile_unpack(ctx->g, vs, ctx->tb, st);
    CK_LAST();
    CK(cudaStreamSynchronize(st));
    NEED(!(ctx->tb.err && *ctx->tb.err == 4), PIQL_ERR_UNSUPPORTED,
         "PIQL_F32_FAST log Z sweeps work on Z = exp(log Z), which left the float range (|log Z| beyond ~87): use PIQL_F32 or PIQL_F64");
    NEED(!(ctx->tb.err && *ctx->tb.err), PIQL_ERR_CUDA, "the persistent sweep kernel timed out waiting for a neighbouring tile");
  }
  CK(cudaEventRecord(ctx->ev1, st));
  if (out) {
    launch_convert_out(vs.x[vs.cur], ctx->d_cell, ctx->d_out, n, vs.precision, st);
    CK_LAST();
    CK(cudaMemcpyAsync(out, ctx->d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  const bool conv = tol > 0 ? residual < tol : true;
  if (stats) {
    stats->sweeps = sweeps; stats->residual = residual; stats->device_seconds = ms * 1e-3;
    stats->backups = sweeps * ctx->nfree; stats->converged = conv ? 1 : 0; stats->kernel = kernel;
  }
  ctx->vi_sweeps_done += sweeps;
  if (!conv) return fail(ctx, PIQL_ERR_NOT_CONVERGED, "max_sweeps reached before the residual fell below tol");
  return PIQL_OK;
}

// Separate log-weight / meta arrays of n entries (the table sampler's records; the split view of d_rec for the host).
int ensure_split(piql_ctx* ctx, long long n) {
  if (n <= ctx->split_cap) return PIQL_OK;
  if (ctx->d_logw) cudaFree(ctx->d_logw);
  if (ctx->d_meta) cudaFree(ctx->d_meta);
  ctx->d_logw = nullptr; ctx->d_meta = nullptr; ctx->split_cap = 0;
  CK(cudaMalloc((void**)&ctx->d_logw, 4 * (size_t)n));
  CK(cudaMalloc((void**)&ctx->d_meta, 4 * (size_t)n));
  ctx->split_cap = n;
  return PIQL_OK;
}

int mc_check_opts(piql_ctx* ctx, const piql_mc_opts* o) {
  NEED(o != nullptr, PIQL_ERR_ARG, "mc opts is null");
  NEED(o->npaths >= 0 && o->path0 >= 0, PIQL_ERR_ARG, "npaths and path0 must be non-negative");
  NEED(o->estimator == 0 || o->estimator == 1, PIQL_ERR_ARG, "estimator must be 0 (sampled stop) or 1 (discounted)");
  NEED(o->horizon > 0 && o->horizon < (1LL << 30), PIQL_ERR_ARG, "horizon must lie in [1, 2^30) (a path record keeps its length in 30 bits)");
  NEED(o->start_mode == PIQL_START_ALL_FREE || o->start_mode == PIQL_START_LIST, PIQL_ERR_ARG, "bad start_mode");
  if (o->start_mode == PIQL_START_LIST) {
    NEED(o->starts != nullptr && o->nstarts > 0, PIQL_ERR_ARG, "start list is empty");
    for (long long i = 0; i < o->nstarts; i++) {
      NEED(o->starts[i] >= 0 && o->starts[i] < ctx->g.n, PIQL_ERR_ARG, "start cell out of range");
      // initial_state never yields a wall (draw_not_wall, GridWorld.jl:54-61), and a wall cell has no policy record
      NEED(!ctx->h_walls_nan(o->starts[i]), PIQL_ERR_ARG, "start cell is a wall");
    }
  }
  if (o->injected) NEED(o->injected_steps > 0, PIQL_ERR_ARG, "injected_steps must be positive");
  return PIQL_OK;
}

// Select the slot list and make sure the accumulators can hold it.
int mc_setup_slots(piql_ctx* ctx, const piql_mc_opts* o) {
  cudaStream_t st = ctx->stream;
  if (o->start_mode == PIQL_START_ALL_FREE) {
    ctx->d_starts = ctx->d_free;
    ctx->nstarts = ctx->nfree;
  } else {
    CK(grow(ctx->d_starts_own, ctx->starts_own_cap, o->nstarts));
    CK(cudaMemcpyAsync(ctx->d_starts_own, o->starts, 4 * (size_t)o->nstarts, cudaMemcpyHostToDevice, st));
    ctx->d_starts = ctx->d_starts_own;
    ctx->nstarts = o->nstarts;
  }
  const long long need = ctx->nstarts * NA;
  if (need > ctx->acc_cap) {
    for (void* p : {(void*)ctx->acc_max, (void*)ctx->acc_sum, (void*)ctx->acc_cnt}) if (p) cudaFree(p);
    ctx->acc_max = ctx->acc_sum = nullptr; ctx->acc_cnt = nullptr; ctx->acc_cap = 0;
    ctx->peer.valid = false;  // exported handles named the old allocation
    CK(cudaMalloc((void**)&ctx->acc_max, 4 * (size_t)need));
    CK(cudaMalloc((void**)&ctx->acc_sum, 4 * (size_t)need));
    CK(cudaMalloc((void**)&ctx->acc_cnt, 4 * (size_t)need));
    ctx->acc_cap = need;
  }
  return PIQL_OK;
}

int mc_reset(piql_ctx* ctx, const piql_mc_opts* o) {
  int rc = mc_setup_slots(ctx, o);
  if (rc) return rc;
  launch_mc_reset(ctx->acc_max, ctx->acc_sum, ctx->acc_cnt, ctx->nstarts * NA, ctx->stream);
  launch_fold_counter(ctx->d_transitions, ctx->stream);  // this run's count moves into the running total
  CK_LAST();
  ctx->paths_done = 0;
  return PIQL_OK;
}

RolloutArgs make_rollout_args(piql_ctx* ctx, const piql_mc_opts* o) {
  RolloutArgs a{};
  a.table = ctx->d_table;
  a.table16 = ctx->t16_valid ? ctx->d_table16 : nullptr;
  a.table_bytes = 32LL * ctx->g.n;
  a.goal = ctx->g.goal; a.H = ctx->g.H;
  a.gamma = ctx->g.gamma; a.rho = ctx->g.randomness; a.vgoal = (float)ctx->v_goal;
  a.lgam = (float)std::log(ctx->g.gamma); a.l1mg = (float)std::log(1.0 - ctx->g.gamma);
  // integer forms of `u >= gamma` / `u < randomness` for u = k * 2^-24 against the problem's Float64 gamma / randomness
  // (products with 2^24 are exact in double): k >= gamma*2^24 <=> k >= ceil(gamma*2^24), k < rho*2^24 <=> k < ceil(rho*2^24).
  // Taking the thresholds from the float-rounded gamma instead would move the sampled stop probability 1 - gamma by up to
  // 2^-25, e.g. 1.00017e-4 instead of 1e-4 at gamma = 0.9999.
  a.gamma_k = (uint32_t)std::ceil(a.gamma * 16777216.0);
  a.rho_k = (uint32_t)std::ceil(a.rho * 16777216.0);
  a.estimator = o->estimator; a.horizon = o->horizon;
  a.starts = ctx->d_starts; a.nstarts = ctx->nstarts;
  a.keys = philox_expand((uint32_t)o->seed, (uint32_t)(o->seed >> 32));
  return a;
}

// Rollout + reduction of opts->npaths paths in launches of at most MAX_RECORDS_PER_LAUNCH.
int mc_rollout(piql_ctx* ctx, const piql_mc_opts* o, int* trace_dev, long long trace_steps) {
  NEED(ctx->has_table, PIQL_ERR_STATE, "no policy table (call piql_policy_table or piql_mc_logz first)");
  cudaStream_t st = ctx->stream;
  const long long chunk = o->npaths < MAX_RECORDS_PER_LAUNCH ? o->npaths : MAX_RECORDS_PER_LAUNCH;
  CK(grow(ctx->d_rec, ctx->rec_cap, chunk));
  ctx->records_split = false;
  const size_t sf = reduce_scratch_floats(chunk, ctx->nstarts);
  CK(grow(ctx->d_scratch, ctx->scratch_cap, sf ? sf : 1));
  const float* inj_dev = nullptr;
  if (o->injected) {
    NEED(o->npaths <= MAX_RECORDS_PER_LAUNCH, PIQL_ERR_ARG, "too many paths for injected uniforms");
    const size_t nf = (size_t)o->npaths * (size_t)o->injected_steps * 4;
    CK(grow(ctx->d_inj, ctx->inj_cap, nf));
    CK(cudaMemcpyAsync(ctx->d_inj, o->injected, 4 * nf, cudaMemcpyHostToDevice, st));
    inj_dev = ctx->d_inj;
  }
  RolloutArgs a = make_rollout_args(ctx, o);
  a.injected = inj_dev; a.injected_steps = o->injected_steps;
  a.rec = ctx->d_rec;
  a.trace = trace_dev; a.trace_steps = trace_steps;
  for (long long done = 0; done < o->npaths; done += chunk) {
    a.path0 = o->path0 + done;
    a.npaths = o->npaths - done < chunk ? o->npaths - done : chunk;
    const bool timed = ctx->tev_used + 3 <= MAX_TIMING_EVENTS;
    if (timed) {
      while (ctx->tev.size() < ctx->tev_used + 3) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->tev.push_back(e);
      }
      CK(cudaEventRecord(ctx->tev[ctx->tev_used], st));
    }
    launch_rollout(a, ctx->num_sms, st);
    if (timed) CK(cudaEventRecord(ctx->tev[ctx->tev_used + 1], st));
    launch_reduce(ctx->d_rec, a.path0, a.npaths, ctx->nstarts, ctx->acc_max, ctx->acc_sum, ctx->acc_cnt,
                  ctx->d_scratch, ctx->scratch_cap, ctx->d_transitions, ctx->num_sms, st);
    if (timed) {
      CK(cudaEventRecord(ctx->tev[ctx->tev_used + 2], st));
      ctx->tev_used += 3;
    }
  }
  CK_LAST();
  ctx->paths_done += (unsigned long long)o->npaths;
  return PIQL_OK;
}

// Sum of the rollout / reduction kernel durations recorded since the last reset (stream must be idle).
int mc_timing(piql_ctx* ctx, double* rollout_s, double* reduce_s, int64_t* launches, bool consume) {
  double ro = 0.0, re = 0.0;
  for (size_t i = 0; i + 3 <= ctx->tev_used; i += 3) {
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, ctx->tev[i], ctx->tev[i + 1]));
    CK(cudaEventElapsedTime(&b, ctx->tev[i + 1], ctx->tev[i + 2]));
    ro += a * 1e-3; re += b * 1e-3;
  }
  if (rollout_s) *rollout_s = ro;
  if (reduce_s) *reduce_s = re;
  if (launches) *launches = (int64_t)(ctx->tev_used / 3);
  if (consume) ctx->tev_used = 0;
  return PIQL_OK;
}

int mc_finalize(piql_ctx* ctx, double* logz_sa, uint64_t* count_sa, double* logz_s, piql_mc_stats* stats) {
  cudaStream_t st = ctx->stream;
  const long long n = ctx->g.n;
  NEED(ctx->nstarts > 0, PIQL_ERR_STATE, "no Monte-Carlo run to finalise");
  // device staging: reuse the VI double buffers (d_in4: 4n doubles, d_out: n doubles)
  int rc = ensure_vi(ctx);
  if (rc) return rc;
  double* d_sa = logz_sa ? ctx->d_in4 : nullptr;
  double* d_s = logz_s ? ctx->d_out : nullptr;
  unsigned long long* d_cnt = count_sa ? reinterpret_cast<unsigned long long*>(ctx->vs.aux4) : nullptr;
  ctx->vs_ready = false;  // aux4 is clobbered
  if (d_sa) launch_fill_f64(d_sa, NAN, 4 * n, st);
  if (d_s) launch_fill_f64(d_s, NAN, n, st);
  if (d_cnt) CK(cudaMemsetAsync(d_cnt, 0, 32 * (size_t)n, st));
  launch_mc_finalize(ctx->acc_max, ctx->acc_sum, ctx->acc_cnt, ctx->d_starts, ctx->nstarts, d_sa, d_cnt, d_s, st);
  CK_LAST();
  if (d_sa) CK(cudaMemcpyAsync(logz_sa, d_sa, 32 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (d_cnt) CK(cudaMemcpyAsync(count_sa, d_cnt, 32 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (d_s) CK(cudaMemcpyAsync(logz_s, d_s, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ctx->h_transitions, ctx->d_transitions, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (ctx->peer.err && *ctx->peer.err)
    return fail(ctx, PIQL_ERR_CUDA, "the peer exchange timed out waiting for another rank (did every rank call it?)");
  if (stats) {
    stats->transitions = *ctx->h_transitions;
    stats->paths = ctx->paths_done;
    double ro = 0.0, re = 0.0;
    int rc2 = mc_timing(ctx, &ro, &re, nullptr, false);
    if (rc2) return rc2;
    stats->rollout_seconds = ro;
    stats->device_seconds = ro + re;
  }
  return PIQL_OK;
}

// Snapshot + second-stream finalisation.  f32: outputs are float / uint32 instead of double / uint64.
int mc_finalize_async(piql_ctx* ctx, void* logz_sa, void* count_sa, void* logz_s, bool f32) {
  cudaStream_t st = ctx->stream;
  const long long n = ctx->g.n, na = ctx->nstarts * NA;
  NEED(ctx->nstarts > 0, PIQL_ERR_STATE, "no Monte-Carlo run to finalise");
  if (!ctx->cstream) {
    CK(cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_snap, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
    CK(cudaMallocHost((void**)&ctx->h_fin_transitions, 16));
  }
  // the previous snapshot must have been consumed before it is overwritten (and before its host buffers are reused)
  if (ctx->fin_pending) CK(cudaStreamWaitEvent(st, ctx->ev_done, 0));
  if (na > ctx->snap_cap) {
    CK(cudaStreamSynchronize(ctx->cstream));
    for (void* p : {(void*)ctx->snap_max, (void*)ctx->snap_sum, (void*)ctx->snap_cnt}) if (p) cudaFree(p);
    ctx->snap_max = ctx->snap_sum = nullptr; ctx->snap_cnt = nullptr; ctx->snap_cap = 0;
    CK(cudaMalloc((void**)&ctx->snap_max, 4 * (size_t)na));
    CK(cudaMalloc((void**)&ctx->snap_sum, 4 * (size_t)na));
    CK(cudaMalloc((void**)&ctx->snap_cnt, 4 * (size_t)na));
    ctx->snap_cap = na;
  }
  const size_t ef = f32 ? 4 : 8;   // bytes per output element (values and counts alike)
  const size_t need = ef * 9 * (size_t)n + 64;
  if (need > ctx->fin_cap) {
    CK(cudaStreamSynchronize(ctx->cstream));
    if (ctx->fin_out) cudaFree(ctx->fin_out);
    ctx->fin_out = nullptr; ctx->fin_cap = 0;
    CK(cudaMalloc((void**)&ctx->fin_out, need));
    ctx->fin_cap = need;
  }
  CK(cudaMemcpyAsync(ctx->snap_max, ctx->acc_max, 4 * (size_t)na, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(ctx->snap_sum, ctx->acc_sum, 4 * (size_t)na, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(ctx->snap_cnt, ctx->acc_cnt, 4 * (size_t)na, cudaMemcpyDeviceToDevice, st));
  unsigned long long* d_tr = reinterpret_cast<unsigned long long*>(ctx->fin_out + ef * 9 * (size_t)n);
  CK(cudaMemcpyAsync(d_tr, ctx->d_transitions, 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaEventRecord(ctx->ev_snap, st));
  cudaStream_t cs = ctx->cstream;
  CK(cudaStreamWaitEvent(cs, ctx->ev_snap, 0));
  unsigned char* d_sa = ctx->fin_out;                      // [n][4] values
  unsigned char* d_cnt = d_sa + ef * 4 * (size_t)n;        // [n][4] counts
  unsigned char* d_s = d_cnt + ef * 4 * (size_t)n;         // [n] values
  if (f32) {
    if (logz_sa) launch_fill_f32((float*)d_sa, NAN, 4 * n, cs);
    if (logz_s) launch_fill_f32((float*)d_s, NAN, n, cs);
    if (count_sa) CK(cudaMemsetAsync(d_cnt, 0, 16 * (size_t)n, cs));
    launch_mc_finalize_f32(ctx->snap_max, ctx->snap_sum, ctx->snap_cnt, ctx->d_starts, ctx->nstarts, logz_sa ? (float*)d_sa : nullptr,
                           count_sa ? (uint32_t*)d_cnt : nullptr, logz_s ? (float*)d_s : nullptr, cs);
  } else {
    if (logz_sa) launch_fill_f64((double*)d_sa, NAN, 4 * n, cs);
    if (logz_s) launch_fill_f64((double*)d_s, NAN, n, cs);
    if (count_sa) CK(cudaMemsetAsync(d_cnt, 0, 32 * (size_t)n, cs));
    launch_mc_finalize(ctx->snap_max, ctx->snap_sum, ctx->snap_cnt, ctx->d_starts, ctx->nstarts, logz_sa ? (double*)d_sa : nullptr,
                       count_sa ? (unsigned long long*)d_cnt : nullptr, logz_s ? (double*)d_s : nullptr, cs);
  }
  CK_LAST();
  if (logz_sa) CK(cudaMemcpyAsync(logz_sa, d_sa, ef * 4 * (size_t)n, cudaMemcpyDeviceToHost, cs));
  if (count_sa) CK(cudaMemcpyAsync(count_sa, d_cnt, ef * 4 * (size_t)n, cudaMemcpyDeviceToHost, cs));
  if (logz_s) CK(cudaMemcpyAsync(logz_s, d_s, ef * (size_t)n, cudaMemcpyDeviceToHost, cs));
  CK(cudaMemcpyAsync(ctx->h_fin_transitions, d_tr, 8, cudaMemcpyDeviceToHost, cs));
  CK(cudaEventRecord(ctx->ev_done, cs));
  ctx->fin_pending = true;
  ctx->fin_paths = ctx->paths_done;
  return PIQL_OK;
}

int build_table(piql_ctx* ctx, const double* V, const double* Q, double beta, float* kl_dev = nullptr) {
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  NEED(V != nullptr, PIQL_ERR_ARG, "V is null");
  NEED(beta == beta, PIQL_ERR_ARG, "beta is NaN");
  int rc = ensure_vi(ctx);
  if (rc) return rc;
  const int n = ctx->g.n;
  cudaStream_t st = ctx->stream;
  const double vg = V[ctx->g.goal];
  NEED(vg == vg, PIQL_ERR_ARG, "V(goal) is NaN");
  CK(cudaMemcpyAsync(ctx->d_in0, V, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  if (Q) CK(cudaMemcpyAsync(ctx->d_in4, Q, 32 * (size_t)n, cudaMemcpyHostToDevice, st));
  ctx->t16_valid = ctx->d_table16 != nullptr && Q == nullptr;
  launch_policy_table(ctx->g, ctx->d_cell, ctx->d_in0, Q ? ctx->d_in4 : nullptr, beta, 1.0, ctx->d_table, st, kl_dev,
                      ctx->t16_valid ? ctx->d_table16 : nullptr);
  CK_LAST();
  ctx->v_goal = vg;
  ctx->has_table = true;
  return PIQL_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int32_t piql_version(void) { return PIQL_VERSION; }

int32_t piql_device_count(int32_t* n_out) {
  if (!n_out) return PIQL_ERR_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { *n_out = 0; return PIQL_ERR_NO_DEVICE; }
  *n_out = n;
  return PIQL_OK;
}

int32_t piql_create(piql_ctx** ctx_out, int32_t device_id) {
  if (!ctx_out) return PIQL_ERR_ARG;
  *ctx_out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return PIQL_ERR_NO_DEVICE;  // no CPU fallback
  if (device_id < 0 || device_id >= n) return PIQL_ERR_ARG;
  piql_ctx* ctx = new (std::nothrow) piql_ctx();
  if (!ctx) return PIQL_ERR_ARG;
  ctx->device = device_id;
  cudaDeviceProp prop;
  bool ok = cudaSetDevice(device_id) == cudaSuccess && cudaGetDeviceProperties(&prop, device_id) == cudaSuccess;
  if (ok && prop.major != 10) {  // built for sm_100a only
    delete ctx;
    return PIQL_ERR_UNSUPPORTED;
  }
  ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
       cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
       cudaMalloc((void**)&ctx->d_scalars, 64) == cudaSuccess &&
       cudaMallocHost((void**)&ctx->h_scalars, 64) == cudaSuccess &&
       cudaMalloc((void**)&ctx->d_transitions, 16) == cudaSuccess &&
       cudaMemset(ctx->d_transitions, 0, 16) == cudaSuccess &&
       cudaMallocHost((void**)&ctx->h_transitions, 16) == cudaSuccess && configure_backup_kernels() == 0;
  if (!ok) {
    cudaGetLastError();
    delete ctx;
    return PIQL_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  *ctx_out = ctx;
  return PIQL_OK;
}

int32_t piql_destroy(piql_ctx* ctx) {
  if (!ctx) return PIQL_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  peer_close(ctx);
  vpeer_close(ctx);
  if (ctx->vpeer.own_flags) cudaFree(ctx->vpeer.own_flags);
  if (ctx->vpeer.wait_stats) cudaFree(ctx->vpeer.wait_stats);
  free_table(ctx);
  if (ctx->peer.own_flags) cudaFree(ctx->peer.own_flags);
  if (ctx->peer.done) cudaFree(ctx->peer.done);
  if (ctx->peer.err) cudaFreeHost(ctx->peer.err);
  free_vi(ctx);
  for (void* p : {(void*)ctx->d_walls, (void*)ctx->d_cell, (void*)ctx->d_free, (void*)ctx->d_table, (void*)ctx->d_rec, (void*)ctx->d_logw,
                  (void*)ctx->d_meta, (void*)ctx->d_starts_own, (void*)ctx->acc_max, (void*)ctx->acc_sum,
                  (void*)ctx->acc_cnt, (void*)ctx->d_scratch, (void*)ctx->d_transitions, (void*)ctx->d_inj,
                  (void*)ctx->d_scalars, (void*)ctx->snap_max, (void*)ctx->snap_sum, (void*)ctx->snap_cnt, (void*)ctx->fin_out})
    if (p) cudaFree(p);
  for (void* p : {(void*)ctx->d_table_spare, (void*)ctx->d_upV, (void*)ctx->d_upQ, (void*)ctx->d_table16, (void*)ctx->d_table16_spare})
    if (p) cudaFree(p);
  for (cudaEvent_t e : {ctx->ev_retire_cur, ctx->ev_retire_spare, ctx->ev_table_ready}) if (e) cudaEventDestroy(e);
  if (ctx->ustream) cudaStreamDestroy(ctx->ustream);
  if (ctx->h_fin_transitions) cudaFreeHost(ctx->h_fin_transitions);
  if (ctx->ev_snap) cudaEventDestroy(ctx->ev_snap);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  if (ctx->cstream) cudaStreamDestroy(ctx->cstream);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  if (ctx->h_transitions) cudaFreeHost(ctx->h_transitions);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->tev) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PIQL_OK;
}

const char* piql_last_error(const piql_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t piql_stream(piql_ctx* ctx, void** stream_out) {
  if (!ctx || !stream_out) return PIQL_ERR_ARG;
  *stream_out = (void*)ctx->stream;
  return PIQL_OK;
}

int32_t piql_launch_count(piql_ctx* ctx, int64_t* n_out) {
  if (!ctx || !n_out) return PIQL_ERR_ARG;
  *n_out = ctx->launches;
  return PIQL_OK;
}

int32_t piql_grid_load(piql_ctx* ctx, const uint8_t* walls, int64_t H, int64_t W, int64_t goal_i, int64_t goal_j,
                       const piql_grid_params* p) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(walls && p, PIQL_ERR_ARG, "walls / params is null");
  NEED(H >= 3 && W >= 3 && H * W < (1LL << 31) / 8, PIQL_ERR_ARG, "grid must be at least 3x3 and below 2^28 cells");
  NEED(goal_i >= 0 && goal_i < H && goal_j >= 0 && goal_j < W, PIQL_ERR_ARG, "goal outside the grid");
  NEED(p->gamma > 0.0 && p->gamma < 1.0, PIQL_ERR_ARG, "gamma must lie in (0, 1)");
  NEED(p->randomness >= 0.0 && p->randomness <= 1.0, PIQL_ERR_ARG, "randomness must lie in [0, 1]");
  const long long n = H * W;
  NEED(!walls[goal_i + H * goal_j], PIQL_ERR_ARG, "goal is a wall");
  // The border must be wall (make_walls, GridWorld.jl:40-45): take_step never leaves the array.
  for (int64_t i = 0; i < H; i++) NEED(walls[i] && walls[i + H * (W - 1)], PIQL_ERR_ARG, "border cells must be walls");
  for (int64_t j = 0; j < W; j++) NEED(walls[H * j] && walls[H - 1 + H * j], PIQL_ERR_ARG, "border cells must be walls");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CK(cudaStreamSynchronize(st));
  if (ctx->ustream) CK(cudaStreamSynchronize(ctx->ustream));
  for (void* q : {(void*)ctx->d_walls, (void*)ctx->d_cell, (void*)ctx->d_free, (void*)ctx->d_table, (void*)ctx->d_table_spare,
                  (void*)ctx->d_table16, (void*)ctx->d_table16_spare})
    if (q) cudaFree(q);
  ctx->d_walls = ctx->d_cell = nullptr; ctx->d_free = nullptr; ctx->d_table = nullptr; ctx->d_table_spare = nullptr;
  ctx->d_table16 = ctx->d_table16_spare = nullptr; ctx->t16_valid = ctx->t16_valid_spare = false;
  ctx->spare_retired = false;
  ctx->has_grid = ctx->has_table = false; ctx->vs_ready = false;
  // free-cell list, in index order (host: it is the tabulation step of the boundary)
  long long nfree = 0;
  for (long long s = 0; s < n; s++) nfree += walls[s] == 0;
  int* h_free = (int*)malloc(sizeof(int) * (size_t)(nfree ? nfree : 1));
  NEED(h_free != nullptr, PIQL_ERR_ARG, "out of host memory");
  long long k = 0;
  for (long long s = 0; s < n; s++) if (!walls[s]) h_free[k++] = (int)s;
  cudaError_t e = cudaMalloc((void**)&ctx->d_walls, (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_cell, (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_free, sizeof(int) * (size_t)(nfree ? nfree : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_table, 32 * (size_t)n);
  {
    // 16-byte records beside the 32-byte table on grids whose table does not stay in L2 (PIQL_ROLL_REC16 = 0 / 1 overrides)
    const char* env16 = getenv("PIQL_ROLL_REC16");
    const bool rec16 = env16 ? atoi(env16) != 0 : n >= ROLL_REC16_MIN_CELLS;
    if (e == cudaSuccess && rec16) e = cudaMalloc((void**)&ctx->d_table16, 16 * (size_t)n);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_walls, walls, (size_t)n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_free, h_free, sizeof(int) * (size_t)nfree, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    ctx->g.H = (int)H; ctx->g.W = (int)W; ctx->g.n = (int)n; ctx->g.goal = (int)(goal_i + H * goal_j);
    ctx->g.randomness = p->randomness; ctx->g.step_cost = p->step_cost; ctx->g.reward = p->reward; ctx->g.gamma = p->gamma;
    ctx->g.bi0 = (int)H; ctx->g.bi1 = -1; ctx->g.bj0 = (int)W; ctx->g.bj1 = -1;
    for (long long q = 0; q < nfree; q++) {
      const int i = h_free[q] % (int)H, j = h_free[q] / (int)H;
      ctx->g.bi0 = i < ctx->g.bi0 ? i : ctx->g.bi0; ctx->g.bi1 = i > ctx->g.bi1 ? i : ctx->g.bi1;
      ctx->g.bj0 = j < ctx->g.bj0 ? j : ctx->g.bj0; ctx->g.bj1 = j > ctx->g.bj1 ? j : ctx->g.bj1;
    }
    launch_cellinfo(ctx->d_walls, ctx->g, ctx->d_cell, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  free(h_free);
  if (e != cudaSuccess) return fail(ctx, PIQL_ERR_CUDA, std::string("piql_grid_load: ") + cudaGetErrorString(e));
  ctx->h_walls.assign(walls, walls + n);
  ctx->nfree = nfree;
  ctx->has_grid = true;
  return PIQL_OK;
}

int32_t piql_grid_free_count(piql_ctx* ctx, int64_t* nfree_out) {
  if (!ctx || !nfree_out) return PIQL_ERR_ARG;
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  *nfree_out = ctx->nfree;
  return PIQL_OK;
}

static int solve_entry(piql_ctx* ctx, int kind, const double* V, const double* Q, const double* nu, double beta,
                       double reward_scale, const piql_solve_opts* opts, double* out, piql_solve_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(opts != nullptr, PIQL_ERR_ARG, "opts is null");
  NEED(out != nullptr, PIQL_ERR_ARG, "output pointer is null");
  CK(cudaSetDevice(ctx->device));
  int rc = vi_prepare(ctx, kind, V, Q, nu, beta, reward_scale, opts->precision);
  if (rc) return rc;
  return vi_solve(ctx, opts, out, stats);
}

int32_t piql_solve_nu(piql_ctx* ctx, double reward_scale, const piql_solve_opts* opts, double* nu_out,
                      piql_solve_stats* stats) {
  return solve_entry(ctx, KIND_NU, nullptr, nullptr, nullptr, 1.0, reward_scale, opts, nu_out, stats);
}
int32_t piql_solve_logz(piql_ctx* ctx, const double* V, const double* Q, double beta, const piql_solve_opts* opts,
                        double* logz_out, piql_solve_stats* stats) {
  return solve_entry(ctx, KIND_LOGZ, V, Q, nullptr, beta, 1.0, opts, logz_out, stats);
}
int32_t piql_solve_gap(piql_ctx* ctx, const double* V, const double* Q, const double* nu, const piql_solve_opts* opts,
                       double* gap_out, piql_solve_stats* stats) {
  return solve_entry(ctx, KIND_GAP, V, Q, nu, 1.0, 1.0, opts, gap_out, stats);
}
int32_t piql_solve_z(piql_ctx* ctx, const double* V, const double* Q, double beta, const piql_solve_opts* opts,
                     double* z_out, piql_solve_stats* stats) {
  return solve_entry(ctx, KIND_Z, V, Q, nullptr, beta, 1.0, opts, z_out, stats);
}

static int update_v_entry(piql_ctx* ctx, int kind, const double* V, double beta, const piql_solve_opts* opts, double* vnew_out,
                          double* logz_out, piql_solve_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(opts != nullptr && vnew_out != nullptr, PIQL_ERR_ARG, "opts / output pointer is null");
  NEED(beta != 0.0, PIQL_ERR_ARG, "beta must be non-zero");
  CK(cudaSetDevice(ctx->device));
  int rc = vi_prepare(ctx, kind, V, nullptr, nullptr, beta, 1.0, opts->precision);
  if (rc) return rc;
  rc = vi_solve(ctx, opts, logz_out, stats);
  if (rc && rc != PIQL_ERR_NOT_CONVERGED) return rc;
  // V + logz/beta   value_iteration.jl:119-124, or V + log(Z)/beta  :53-59 (d_in0 still holds the caller's V in double)
  cudaStream_t st = ctx->stream;
  launch_update_v(ctx->d_in0, ctx->vs.x[ctx->vs.cur], ctx->d_cell, beta, ctx->d_out, ctx->g.n, ctx->vs.precision, st,
                  kind == KIND_Z);
  CK_LAST();
  CK(cudaMemcpyAsync(vnew_out, ctx->d_out, 8 * (size_t)ctx->g.n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return rc;
}
int32_t piql_update_v(piql_ctx* ctx, const double* V, double beta, const piql_solve_opts* opts, double* vnew_out,
                      double* logz_out, piql_solve_stats* stats) {
  return update_v_entry(ctx, KIND_LOGZ, V, beta, opts, vnew_out, logz_out, stats);
}
int32_t piql_z_update_v(piql_ctx* ctx, const double* V, double beta, const piql_solve_opts* opts, double* vnew_out,
                        double* z_out, piql_solve_stats* stats) {
  return update_v_entry(ctx, KIND_Z, V, beta, opts, vnew_out, z_out, stats);
}

// ---- batched entry points: the experiment protocol in two launches ------------------------------------------
static int batch_run(piql_ctx* ctx, int kind, const double* Vin, long long vin_count, const double* nu, const double* betas,
                     int nproblems, int nupdates, const piql_solve_opts* opts, double* out, piql_solve_stats* stats_out) {
  const int n = ctx->g.n;
  cudaStream_t st = ctx->stream;
  const int prec = opts->precision == PIQL_F64 ? 0 : 1;
  const int nsolves = kind == KIND_LOGZ ? nupdates : 1;
  const size_t out_per = (kind == KIND_LOGZ ? (size_t)(nupdates + 1) : 1) * (size_t)n;
  double *d_v = nullptr, *d_nu = nullptr, *d_b = nullptr, *d_o = nullptr, *d_s = nullptr;
  std::vector<double> h_s((size_t)nproblems * nsolves * 2);
  cudaError_t e = cudaMalloc((void**)&d_v, 8 * (size_t)n * (size_t)vin_count);
  if (e == cudaSuccess && nu) e = cudaMalloc((void**)&d_nu, 8 * (size_t)n);
  if (e == cudaSuccess && betas) e = cudaMalloc((void**)&d_b, 8 * (size_t)nproblems);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_o, 8 * out_per * nproblems);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_s, 8 * h_s.size());
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, Vin, 8 * (size_t)n * (size_t)vin_count, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && nu) e = cudaMemcpyAsync(d_nu, nu, 8 * (size_t)n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && betas) e = cudaMemcpyAsync(d_b, betas, 8 * (size_t)nproblems, cudaMemcpyHostToDevice, st);
  float ms = 0.f;
  if (e == cudaSuccess) {
    cudaEventRecord(ctx->ev0, st);
    launch_resident_batch(ctx->g, ctx->d_cell, kind, prec, 1.0, nproblems, d_v, vin_count > 1 ? n : 0, d_nu, d_b, nupdates,
                          opts->tol > 0 ? opts->tol : -1.0, opts->max_sweeps, d_o, d_s, st);
    e = cudaGetLastError();
    cudaEventRecord(ctx->ev1, st);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, 8 * out_per * nproblems, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_s.data(), d_s, 8 * h_s.size(), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  for (void* p : {(void*)d_v, (void*)d_nu, (void*)d_b, (void*)d_o, (void*)d_s}) if (p) cudaFree(p);
  if (e != cudaSuccess) return fail(ctx, PIQL_ERR_CUDA, std::string("batched solve: ") + cudaGetErrorString(e));
  bool all = true;
  for (size_t i = 0; i < h_s.size() / 2; i++) {
    const bool conv = opts->tol > 0 ? h_s[2 * i + 1] < opts->tol : true;
    all = all && conv;
    if (stats_out) {
      piql_solve_stats& o = stats_out[i];
      o.sweeps = (int64_t)h_s[2 * i]; o.residual = h_s[2 * i + 1]; o.device_seconds = ms * 1e-3;  // of the whole batch
      o.backups = o.sweeps * ctx->nfree; o.converged = conv ? 1 : 0; o.kernel = 3;
    }
  }
  if (!all) return fail(ctx, PIQL_ERR_NOT_CONVERGED, "max_sweeps reached before the residual fell below tol in at least one solve");
  return PIQL_OK;
}

int32_t piql_learning_traj(piql_ctx* ctx, const double* V0, const double* betas, int32_t nbeta, int32_t nupdates,
                           const piql_solve_opts* opts, double* traj_out, piql_solve_stats* stats_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  NEED(V0 && betas && opts && traj_out, PIQL_ERR_ARG, "V0 / betas / opts / traj_out is null");
  NEED(nbeta > 0 && nupdates > 0, PIQL_ERR_ARG, "nbeta and nupdates must be positive");
  NEED(opts->max_sweeps > 0, PIQL_ERR_ARG, "max_sweeps must be positive");
  NEED(opts->precision >= PIQL_F64 && opts->precision <= PIQL_F32_FAST, PIQL_ERR_ARG, "bad precision");
  for (int b = 0; b < nbeta; b++) NEED(betas[b] == betas[b] && betas[b] != 0.0, PIQL_ERR_ARG, "beta must be non-zero and not NaN");
  NEED(V0[ctx->g.goal] == V0[ctx->g.goal], PIQL_ERR_ARG, "V(goal) is NaN");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->g.n;
  if (resident_batch_smem_bytes(n, opts->precision == PIQL_F64 ? 0 : 1) <= 200 * 1024)
    return batch_run(ctx, KIND_LOGZ, V0, 1, nullptr, betas, nbeta, nupdates, opts, traj_out, stats_out);
  // large grids: the chains run one after the other on the whole GPU (tiled kernel)
  int worst = PIQL_OK;
  for (int b = 0; b < nbeta; b++) {
    double* tb = traj_out + (size_t)b * (nupdates + 1) * n;
    for (int s = 0; s < n; s++) tb[s] = ctx->h_walls_nan(s) ? NAN : V0[s];
    for (int u = 0; u < nupdates; u++) {
      std::vector<double> vin(tb + (size_t)u * n, tb + (size_t)(u + 1) * n);
      for (double& v : vin) if (!(v == v)) v = 0.0;
      int rc = piql_update_v(ctx, vin.data(), betas[b], opts, tb + (size_t)(u + 1) * n, nullptr,
                             stats_out ? stats_out + (size_t)b * nupdates + u : nullptr);
      if (rc && rc != PIQL_ERR_NOT_CONVERGED) return rc;
      if (rc) worst = rc;
    }
  }
  return worst;
}

int32_t piql_solve_gap_batch(piql_ctx* ctx, const double* Vs, int32_t m, const double* nu, const piql_solve_opts* opts,
                             double* gaps_out, piql_solve_stats* stats_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  NEED(Vs && nu && opts && gaps_out, PIQL_ERR_ARG, "Vs / nu / opts / gaps_out is null");
  NEED(m > 0, PIQL_ERR_ARG, "m must be positive");
  NEED(opts->max_sweeps > 0, PIQL_ERR_ARG, "max_sweeps must be positive");
  NEED(opts->precision >= PIQL_F64 && opts->precision <= PIQL_F32_FAST, PIQL_ERR_ARG, "bad precision");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->g.n;
  if (resident_batch_smem_bytes(n, opts->precision == PIQL_F64 ? 0 : 1) <= 200 * 1024)
    return batch_run(ctx, KIND_GAP, Vs, m, nu, nullptr, m, 1, opts, gaps_out, stats_out);
  int worst = PIQL_OK;
  for (int b = 0; b < m; b++) {
    int rc = piql_solve_gap(ctx, Vs + (size_t)b * n, nullptr, nu, opts, gaps_out + (size_t)b * n, stats_out ? stats_out + b : nullptr);
    if (rc && rc != PIQL_ERR_NOT_CONVERGED) return rc;
    if (rc) worst = rc;
  }
  return worst;
}

// ---- tabulated ControlProblems ---------------------------------------------------------------------------------
int32_t piql_table_load(piql_ctx* ctx, const piql_table_problem* p) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(p != nullptr, PIQL_ERR_ARG, "problem is null");
  NEED(p->nstates > 0 && p->nstates < (1LL << 31), PIQL_ERR_ARG, "nstates must lie in [1, 2^31)");
  NEED(p->nactions >= 1 && p->nactions <= 16, PIQL_ERR_ARG, "nactions must lie in [1, 16]");
  NEED(p->row_ptr && p->succ && p->prob && p->reward && p->log_prior && p->terminal, PIQL_ERR_ARG, "a table array is null");
  NEED(p->gamma > 0.0 && p->gamma <= 1.0, PIQL_ERR_ARG, "gamma must lie in (0, 1]");
  const long long rows = p->nstates * p->nactions;
  NEED(p->row_ptr[0] == 0, PIQL_ERR_ARG, "row_ptr[0] must be 0");
  for (long long r = 0; r < rows; r++) NEED(p->row_ptr[r + 1] >= p->row_ptr[r], PIQL_ERR_ARG, "row_ptr must be non-decreasing");
  const long long nnz = p->row_ptr[rows];
  for (long long s = 0; s < p->nstates; s++) {
    for (int a = 0; a < p->nactions; a++) {
      double tot = 0.0;
      for (long long e = p->row_ptr[s * p->nactions + a]; e < p->row_ptr[s * p->nactions + a + 1]; e++) {
        NEED(p->succ[e] >= 0 && p->succ[e] < p->nstates, PIQL_ERR_ARG, "successor index out of range");
        NEED(p->prob[e] >= 0.0, PIQL_ERR_ARG, "negative transition probability");
        tot += p->prob[e];
      }
      NEED(p->terminal[s] || std::fabs(tot - 1.0) < 1e-9, PIQL_ERR_ARG, "transition probabilities of a non-terminal (s, a) must sum to 1");
    }
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CK(cudaStreamSynchronize(st));
  free_table(ctx);
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes ? bytes : 8);
    if (e != cudaSuccess) return e;
    ctx->tab_bufs.push_back(*dst);
    return src ? cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
  };
  void *d_ptr, *d_succ, *d_prob, *d_rew, *d_lp, *d_term;
  const size_t ns = (size_t)p->nstates;
  CK(up(p->row_ptr, 8 * (size_t)(rows + 1), &d_ptr));
  CK(up(p->succ, 4 * (size_t)nnz, &d_succ));
  CK(up(p->prob, 8 * (size_t)nnz, &d_prob));
  CK(up(p->reward, 8 * (size_t)nnz, &d_rew));
  CK(up(p->log_prior, 8 * (size_t)rows, &d_lp));
  CK(up(p->terminal, ns, &d_term));
  CK(up(nullptr, 8 * ns, (void**)&ctx->tb_V));
  CK(up(nullptr, 8 * ns, (void**)&ctx->tb_nu));
  CK(up(nullptr, 8 * ns, (void**)&ctx->tb_x[0]));
  CK(up(nullptr, 8 * ns, (void**)&ctx->tb_x[1]));
  CK(up(nullptr, 8 * (size_t)rows, (void**)&ctx->tb_aux));
  CK(up(nullptr, 8 * (size_t)rows, (void**)&ctx->tb_Q));
  ctx->tb_has_q = false;
  CK(up(nullptr, 8 * ns, (void**)&ctx->tb_aux1));
  CK(up(nullptr, 8 * (size_t)table_blocks(p->nstates), (void**)&ctx->tb_partial));
  CK(up(nullptr, 8 * ns, (void**)&ctx->tb_out));
  CK(cudaStreamSynchronize(st));
  TableDev& t = ctx->tab;
  t.ns = p->nstates; t.na = p->nactions;
  t.ptr = (const long long*)d_ptr; t.succ = (const int*)d_succ; t.prob = (const double*)d_prob; t.rew = (const double*)d_rew;
  t.lprior = (const double*)d_lp; t.term = (const uint8_t*)d_term;
  t.gamma = p->gamma; t.lgamma = std::log(p->gamma); t.l1mgamma = std::log(1.0 - p->gamma);  // -inf at gamma = 1
  ctx->tab_nnz = nnz;
  ctx->has_table_problem = true;
  return PIQL_OK;
}

int32_t piql_table_set_q(piql_ctx* ctx, const double* Q) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(ctx->has_table_problem, PIQL_ERR_STATE, "no tabulated problem loaded (call piql_table_load first)");
  CK(cudaSetDevice(ctx->device));
  if (Q) {
    const size_t rows = (size_t)ctx->tab.ns * (size_t)ctx->tab.na;
    for (size_t i = 0; i < rows; i++) NEED(Q[i] == Q[i], PIQL_ERR_ARG, "Q holds a NaN");
    CK(cudaMemcpyAsync(ctx->tb_Q, Q, 8 * rows, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->tb_has_q = Q != nullptr;
  return PIQL_OK;
}

static int table_solve(piql_ctx* ctx, int kind, const double* V, const double* nu, double beta, double reward_scale,
                       const piql_solve_opts* opts, double* x_out, piql_solve_stats* stats, int* cur_out) {
  NEED(ctx->has_table_problem, PIQL_ERR_STATE, "no tabulated problem loaded (call piql_table_load first)");
  NEED(kind >= 0 && kind <= 2, PIQL_ERR_ARG, "kind must be 0 (nu), 1 (logz) or 2 (gap)");
  NEED(opts != nullptr && opts->max_sweeps > 0, PIQL_ERR_ARG, "opts is null or max_sweeps is not positive");
  NEED(kind == KIND_NU || V != nullptr, PIQL_ERR_ARG, "V is required for logz / gap");
  NEED(kind != KIND_GAP || nu != nullptr, PIQL_ERR_ARG, "nu is required for the optimality gap");
  const TableDev& t = ctx->tab;
  cudaStream_t st = ctx->stream;
  const size_t nb = 8 * (size_t)t.ns;
  if (V) CK(cudaMemcpyAsync(ctx->tb_V, V, nb, cudaMemcpyHostToDevice, st));
  if (nu) CK(cudaMemcpyAsync(ctx->tb_nu, nu, nb, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->tb_x[0], 0, nb, st));  // Growable default x -> 0.0
  if (kind != KIND_NU)
    launch_table_prep(t, kind, reward_scale, beta, ctx->tb_V, ctx->tb_nu, ctx->tb_aux, ctx->tb_aux1, st, ctx->tb_has_q ? ctx->tb_Q : nullptr);
  CK_LAST();
  const long long every = opts->check_every > 0 ? opts->check_every : 16;
  long long sweeps = 0;
  double residual = INFINITY;
  int cur = 0;
  CK(cudaEventRecord(ctx->ev0, st));
  while (sweeps < opts->max_sweeps) {
    const long long batch = std::min<long long>(opts->max_sweeps - sweeps, every);
    for (long long i = 0; i < batch; i++) {
      const bool last = i == batch - 1;
      launch_table_sweep(t, kind, reward_scale, ctx->tb_x[cur], ctx->tb_x[cur ^ 1], ctx->tb_V, ctx->tb_aux, ctx->tb_aux1,
                         last ? ctx->tb_partial : nullptr, ctx->d_scalars, st);
      cur ^= 1;
    }
    CK_LAST();
    sweeps += batch;
    CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    residual = ctx->h_scalars[0];
    if (!(residual == residual)) return fail(ctx, PIQL_ERR_CUDA, "soft-VI residual is NaN (diverged)");
    if (opts->tol > 0 && residual < opts->tol) break;
  }
  CK(cudaEventRecord(ctx->ev1, st));
  if (x_out) CK(cudaMemcpyAsync(x_out, ctx->tb_x[cur], nb, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  const bool conv = opts->tol > 0 ? residual < opts->tol : true;
  if (stats) {
    stats->sweeps = sweeps; stats->residual = residual; stats->device_seconds = ms * 1e-3;
    stats->backups = sweeps * t.ns; stats->converged = conv ? 1 : 0; stats->kernel = 4;
  }
  if (cur_out) *cur_out = cur;
  if (!conv) return fail(ctx, PIQL_ERR_NOT_CONVERGED, "max_sweeps reached before the residual fell below tol");
  return PIQL_OK;
}

int32_t piql_table_solve(piql_ctx* ctx, int32_t kind, const double* V, const double* nu, double beta, double reward_scale,
                         const piql_solve_opts* opts, double* x_out, piql_solve_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(x_out != nullptr, PIQL_ERR_ARG, "output pointer is null");
  CK(cudaSetDevice(ctx->device));
  return table_solve(ctx, kind, V, nu, beta, reward_scale, opts, x_out, stats, nullptr);
}

int32_t piql_table_update_v(piql_ctx* ctx, const double* V, double beta, const piql_solve_opts* opts, double* vnew_out,
                            double* logz_out, piql_solve_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(vnew_out != nullptr, PIQL_ERR_ARG, "output pointer is null");
  NEED(beta == beta && beta != 0.0, PIQL_ERR_ARG, "beta must be non-zero and not NaN");
  CK(cudaSetDevice(ctx->device));
  int cur = 0;
  int rc = table_solve(ctx, KIND_LOGZ, V, nullptr, beta, 1.0, opts, logz_out, stats, &cur);
  if (rc && rc != PIQL_ERR_NOT_CONVERGED) return rc;
  launch_table_update_v(ctx->tab.ns, ctx->tb_V, ctx->tb_x[cur], beta, ctx->tb_out, ctx->stream);
  CK_LAST();
  CK(cudaMemcpyAsync(vnew_out, ctx->tb_out, 8 * (size_t)ctx->tab.ns, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return rc;
}

int32_t piql_table_mc_logz(piql_ctx* ctx, const double* V, double beta, const piql_mc_opts* o, double* logz_sa_out,
                           uint64_t* count_sa_out, double* logz_s_out, piql_mc_stats* stats, float* cdf_out, float* bf_out,
                           float* cum_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->has_table_problem, PIQL_ERR_STATE, "no tabulated problem loaded (call piql_table_load first)");
  NEED(V != nullptr && o != nullptr, PIQL_ERR_ARG, "V / opts is null");
  NEED(beta == beta, PIQL_ERR_ARG, "beta is NaN");
  NEED(o->npaths >= 0 && o->path0 >= 0 && o->horizon > 0, PIQL_ERR_ARG, "npaths / path0 / horizon out of range");
  NEED(o->estimator == PIQL_EST_SAMPLED_STOP, PIQL_ERR_UNSUPPORTED, "tabulated problems support the sampled-stop estimator only");
  NEED(o->injected == nullptr, PIQL_ERR_UNSUPPORTED, "injected uniforms are not supported on tabulated problems");
  NEED(o->npaths <= MAX_RECORDS_PER_LAUNCH, PIQL_ERR_ARG, "too many paths for one call (at most 2^28)");
  const TableDev& t = ctx->tab;
  if (o->start_mode == PIQL_START_LIST) {
    NEED(o->starts != nullptr && o->nstarts > 0, PIQL_ERR_ARG, "start list is empty");
    for (long long i = 0; i < o->nstarts; i++) NEED(o->starts[i] >= 0 && o->starts[i] < t.ns, PIQL_ERR_ARG, "start state out of range");
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long long rows = t.ns * t.na, nstarts = o->start_mode == PIQL_START_LIST ? o->nstarts : t.ns, nb = nstarts * t.na;
  {
    int rc2 = ensure_split(ctx, o->npaths ? o->npaths : 1);
    if (rc2) return rc2;
    ctx->records_split = true;
  }
  std::vector<void*> tmp;
  auto dev = [&](size_t bytes, void** p) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
    if (e == cudaSuccess) tmp.push_back(*p);
    return e;
  };
  float *d_cdf, *d_bf, *d_cum, *d_vt;
  int *d_omax, *d_starts = nullptr;
  double *d_sum, *d_lsa, *d_ls;
  unsigned long long *d_cnt, *d_csa;
  cudaError_t e = dev(4 * (size_t)rows, (void**)&d_cdf);
  if (e == cudaSuccess) e = dev(4 * (size_t)rows, (void**)&d_bf);
  if (e == cudaSuccess) e = dev(4 * (size_t)ctx->tab_nnz, (void**)&d_cum);
  if (e == cudaSuccess) e = dev(4 * (size_t)t.ns, (void**)&d_vt);
  if (e == cudaSuccess) e = dev(4 * (size_t)nb, (void**)&d_omax);
  if (e == cudaSuccess) e = dev(8 * (size_t)nb, (void**)&d_sum);
  if (e == cudaSuccess) e = dev(8 * (size_t)nb, (void**)&d_cnt);
  if (e == cudaSuccess) e = dev(8 * (size_t)rows, (void**)&d_lsa);
  if (e == cudaSuccess) e = dev(8 * (size_t)rows, (void**)&d_csa);
  if (e == cudaSuccess) e = dev(8 * (size_t)t.ns, (void**)&d_ls);
  if (e == cudaSuccess && o->start_mode == PIQL_START_LIST) {
    e = dev(4 * (size_t)nstarts, (void**)&d_starts);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_starts, o->starts, 4 * (size_t)nstarts, cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->tb_V, V, 8 * (size_t)t.ns, cudaMemcpyHostToDevice, st);
  float ms_roll = 0.f, ms_all = 0.f;
  if (e == cudaSuccess) {
    launch_fold_counter(ctx->d_transitions, st);
    launch_table_policy(t, 1.0, beta, ctx->tb_V, ctx->tb_has_q ? ctx->tb_Q : nullptr, d_cdf, d_bf, d_cum, d_vt, st);
    launch_fill_f64(d_lsa, NAN, rows, st);
    launch_fill_f64(d_ls, NAN, t.ns, st);
    e = cudaMemsetAsync(d_csa, 0, 8 * (size_t)rows, st);
    cudaEventRecord(ctx->ev0, st);
    launch_table_rollout(t, d_cdf, d_bf, d_cum, d_vt, t.gamma, o->horizon, o->path0, o->npaths, d_starts, nstarts, o->seed,
                         ctx->d_logw, ctx->d_meta, ctx->num_sms, st);
    cudaEventRecord(ctx->ev1, st);
    launch_table_reduce(ctx->d_logw, ctx->d_meta, o->path0, o->npaths, d_starts, nstarts, t.na, d_omax, d_sum, d_cnt,
                        ctx->d_transitions, d_lsa, d_csa, d_ls, ctx->num_sms, st);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e == cudaSuccess && logz_sa_out) e = cudaMemcpyAsync(logz_sa_out, d_lsa, 8 * (size_t)rows, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && count_sa_out) e = cudaMemcpyAsync(count_sa_out, d_csa, 8 * (size_t)rows, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && logz_s_out) e = cudaMemcpyAsync(logz_s_out, d_ls, 8 * (size_t)t.ns, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && cdf_out) e = cudaMemcpyAsync(cdf_out, d_cdf, 4 * (size_t)rows, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && bf_out) e = cudaMemcpyAsync(bf_out, d_bf, 4 * (size_t)rows, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && cum_out) e = cudaMemcpyAsync(cum_out, d_cum, 4 * (size_t)ctx->tab_nnz, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_transitions, ctx->d_transitions, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) {
    cudaEventElapsedTime(&ms_roll, ctx->ev0, ctx->ev1);
    ms_all = ms_roll;
  }
  for (void* p : tmp) cudaFree(p);
  if (e != cudaSuccess) return fail(ctx, PIQL_ERR_CUDA, std::string("piql_table_mc_logz: ") + cudaGetErrorString(e));
  if (stats) {
    stats->transitions = *ctx->h_transitions; stats->paths = (uint64_t)o->npaths;
    stats->rollout_seconds = ms_roll * 1e-3; stats->device_seconds = ms_all * 1e-3;
  }
  return PIQL_OK;
}

int32_t piql_sweep_once(piql_ctx* ctx, int32_t kind, const double* V, const double* Q, const double* nu, double beta,
                        double reward_scale, int32_t precision, const double* x_in, double* x_out, double* residual_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(x_in && x_out, PIQL_ERR_ARG, "x_in / x_out is null");
  CK(cudaSetDevice(ctx->device));
  int rc = vi_prepare(ctx, kind, V, Q, nu, beta, reward_scale, precision);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const int n = ctx->g.n;
  CK(cudaMemcpyAsync(ctx->d_out, x_in, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_convert_in(ctx->d_out, ctx->vs.x[0], n, precision, st);
  launch_sweep(ctx->g, ctx->d_cell, ctx->vs, true, st);
  launch_convert_out(ctx->vs.x[ctx->vs.cur], ctx->d_cell, ctx->d_out, n, precision, st);
  CK_LAST();
  CK(cudaMemcpyAsync(x_out, ctx->d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (residual_out) *residual_out = ctx->h_scalars[0];
  return PIQL_OK;
}

int32_t piql_vi_prepare_dev(piql_ctx* ctx, int32_t kind, const double* V, const double* Q, const double* nu, double beta,
                            double reward_scale, int32_t precision) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(kind != KIND_Z, PIQL_ERR_ARG, "the device-resident sweep loop covers kinds 0..2 (use piql_solve_z for the linear-space recursion)");
  int rc = vi_prepare(ctx, kind, V, Q, nu, beta, reward_scale, precision);
  if (rc) return rc;
  if (ctx->g.n > RESIDENT_MAX_CELLS) CK(tile_begin(ctx->g, ctx->d_cell, ctx->vs, ctx->tb, ctx->num_sms, ctx->stream));
  return PIQL_OK;
}
int32_t piql_vi_tile_plan(piql_ctx* ctx, int32_t kind, int32_t precision, int32_t* plan_out) {
  if (!ctx || !plan_out) return PIQL_ERR_ARG;
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  NEED(kind >= KIND_NU && kind <= KIND_GAP, PIQL_ERR_ARG, "kind must be 0 (nu), 1 (logz) or 2 (gap)");
  bool persistent = false;
  TilePlan pl = choose_tile_plan(ctx->g.H, ctx->g.W, precision == PIQL_F64 ? 0 : 1, kind, precision == PIQL_F32_FAST, ctx->num_sms, &persistent);
  if (ctx->tb.active && ctx->vs.kind == kind && ctx->vs.precision == (precision == PIQL_F64 ? 0 : 1)) {
    pl = ctx->tb.plan;  // the shape in use (timed on the device), not just the model's choice
    persistent = ctx->tb.persistent;
  }
  plan_out[0] = pl.TJ; plan_out[1] = pl.nw; plan_out[2] = pl.tiles_i; plan_out[3] = pl.tiles_j;
  plan_out[4] = pl.ctas_per_sm; plan_out[5] = (int32_t)pl.smem; plan_out[6] = pl.kh; plan_out[7] = (persistent ? 1 : 0) | (pl.planes << 1) | (pl.colres << 4);
  return PIQL_OK;
}
int32_t piql_vi_set_iterate(piql_ctx* ctx, const double* x) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->vs_ready, PIQL_ERR_STATE, "call piql_vi_prepare_dev first");
  NEED(x != nullptr, PIQL_ERR_ARG, "x is null");
  NEED(!ctx->tb.sharded, PIQL_ERR_STATE, "set the iterate before piql_vi_shard_begin");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = ctx->g.n;
  CK(cudaMemcpyAsync(ctx->d_out, x, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_convert_in(ctx->d_out, ctx->vs.x[ctx->vs.cur], n, ctx->vs.precision, st);
  CK_LAST();
  if (n > RESIDENT_MAX_CELLS) CK(tile_begin(ctx->g, ctx->d_cell, ctx->vs, ctx->tb, ctx->num_sms, st));
  CK(cudaStreamSynchronize(st));  // x is the caller's buffer
  return PIQL_OK;
}
int32_t piql_vi_sweeps_dev(piql_ctx* ctx, int64_t sweeps) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->vs_ready, PIQL_ERR_STATE, "call piql_vi_prepare_dev first");
  NEED(sweeps > 0, PIQL_ERR_ARG, "sweeps must be positive");
  CK(cudaSetDevice(ctx->device));
  if (ctx->tb.active && ctx->tb.sharded) {
    NEED(ctx->vpeer.valid, PIQL_ERR_STATE, "call piql_vi_shard_connect first");
    CK(tile_sweeps_sharded(ctx->g, ctx->vs, ctx->tb, ctx->vpeer, sweeps, true, ctx->stream));
  } else if (ctx->tb.active) {
    CK(tile_sweeps(ctx->g, ctx->vs, ctx->tb, sweeps, true, ctx->stream));
  } else {
    for (int64_t i = 0; i < sweeps; i++) launch_sweep(ctx->g, ctx->d_cell, ctx->vs, i == sweeps - 1, ctx->stream);
    CK_LAST();
  }
  ctx->vi_sweeps_done += sweeps;
  return PIQL_OK;
}
int32_t piql_vi_fetch(piql_ctx* ctx, double* x_out, double* residual_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->vs_ready, PIQL_ERR_STATE, "call piql_vi_prepare_dev first");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = ctx->g.n;
  if (x_out) {
    if (ctx->tb.active) tile_unpack(ctx->g, ctx->vs, ctx->tb, st);
    launch_convert_out(ctx->vs.x[ctx->vs.cur], ctx->d_cell, ctx->d_out, n, ctx->vs.precision, st);
    CK_LAST();
    CK(cudaMemcpyAsync(x_out, ctx->d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  NEED(!(ctx->tb.err && *ctx->tb.err == 4), PIQL_ERR_UNSUPPORTED,
       "PIQL_F32_FAST log Z sweeps work on Z = exp(log Z), which left the float range (|log Z| beyond ~87): use PIQL_F32 or PIQL_F64");
  NEED(!(ctx->tb.err && *ctx->tb.err), PIQL_ERR_CUDA, "a soft-VI kernel timed out waiting for a neighbouring tile or rank");
  if (residual_out) *residual_out = ctx->h_scalars[0];
  return PIQL_OK;
}

int32_t piql_vi_shard_begin(piql_ctx* ctx, int32_t rank, int32_t world, piql_vi_shard_handles* out) {
  if (!ctx || !out) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->vs_ready, PIQL_ERR_STATE, "call piql_vi_prepare_dev first");
  NEED(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, PIQL_ERR_ARG, "rank / world out of range (at most 8 GPUs)");
  NEED(ctx->g.n > RESIDENT_MAX_CELLS, PIQL_ERR_UNSUPPORTED, "grids of at most 8192 cells are solved by one CTA; nothing to shard");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  vpeer_close(ctx);
  ViPeer& vp = ctx->vpeer;
  ctx->tb.sharded = true;
  // the shape every rank uses, then this rank's slab of tile columns
  const TilePlan pl = make_tile_plan(ctx->g.H, ctx->g.W, ctx->vs.precision, ctx->vs.kind, ctx->vs.fast, ctx->num_sms);
  NEED(pl.tiles_j >= world, PIQL_ERR_UNSUPPORTED, "fewer tile columns than ranks");
  ctx->tb.tj0 = (int)((long long)pl.tiles_j * rank / world);
  ctx->tb.tj1 = (int)((long long)pl.tiles_j * (rank + 1) / world);
  CK(tile_begin(ctx->g, ctx->d_cell, ctx->vs, ctx->tb, ctx->num_sms, st));  // packs the iterate in the shared shape
  if (!vp.own_flags) CK(cudaMalloc((void**)&vp.own_flags, 4 * sizeof(unsigned)));
  if (!vp.wait_stats) CK(cudaMalloc((void**)&vp.wait_stats, 2 * sizeof(unsigned long long)));
  CK(cudaStreamSynchronize(st));
  CK(cudaMemset(vp.own_flags, 0, 4 * sizeof(unsigned)));
  CK(cudaMemset(vp.wait_stats, 0, 2 * sizeof(unsigned long long)));
  vp.done = vp.own_flags + 2;
  vp.round = 0; vp.rank = rank; vp.world = world;
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->x0, ctx->tb.x[0]));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->x1, ctx->tb.x[1]));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->flags, vp.own_flags));
  out->np = ctx->tb.plan.np; out->tj = ctx->tb.plan.TJ; out->tiles_j = ctx->tb.plan.tiles_j;
  // real columns this rank owns: [col0, col1)
  out->col0 = (int64_t)ctx->tb.tj0 * pl.TJ;
  out->col1 = std::min<int64_t>((int64_t)ctx->tb.tj1 * pl.TJ, ctx->g.W);
  return PIQL_OK;
}
int32_t piql_vi_shard_connect(piql_ctx* ctx, const piql_vi_shard_handles* all) {
  if (!ctx || !all) return PIQL_ERR_ARG;
  ViPeer& vp = ctx->vpeer;
  NEED(ctx->tb.sharded && vp.own_flags, PIQL_ERR_STATE, "call piql_vi_shard_begin first");
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < vp.world; r++)
    NEED(all[r].np == ctx->tb.plan.np && all[r].tj == ctx->tb.plan.TJ, PIQL_ERR_ARG, "ranks disagree on the tile shape (different grids?)");
  for (int r = 0; r < vp.world; r++) {
    if (r == vp.rank) { vp.x[0][r] = ctx->tb.x[0]; vp.x[1][r] = ctx->tb.x[1]; vp.flags[r] = vp.own_flags; continue; }
    if (r != vp.rank - 1 && r != vp.rank + 1) continue;  // only the neighbours are ever touched
    void* p[3] = {nullptr, nullptr, nullptr};
    const uint8_t* h[3] = {all[r].x0, all[r].x1, all[r].flags};
    for (int q = 0; q < 3; q++) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, h[q], sizeof(hd));
      cudaError_t e = cudaIpcOpenMemHandle(&p[q], hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        vpeer_close(ctx);
        return fail(ctx, PIQL_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      }
      vp.opened[vp.nopened++] = p[q];
    }
    vp.x[0][r] = p[0]; vp.x[1][r] = p[1]; vp.flags[r] = (unsigned*)p[2];
  }
  vp.valid = true;
  return PIQL_OK;
}
int32_t piql_vi_shard_wait_stats(piql_ctx* ctx, double* wait_us_out, int64_t* waits_out) {
  if (!ctx) return PIQL_ERR_ARG;
  ViPeer& vp = ctx->vpeer;
  NEED(vp.wait_stats != nullptr, PIQL_ERR_STATE, "call piql_vi_shard_begin first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpy(h, vp.wait_stats, sizeof(h), cudaMemcpyDeviceToHost));
  CK(cudaMemset(vp.wait_stats, 0, sizeof(h)));
  if (wait_us_out) *wait_us_out = 1e-3 * (double)h[0];
  if (waits_out) *waits_out = (int64_t)h[1];
  return PIQL_OK;
}
int32_t piql_vi_shard_end(piql_ctx* ctx) {
  if (!ctx) return PIQL_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  vpeer_close(ctx);
  return PIQL_OK;
}

int32_t piql_policy_table(piql_ctx* ctx, const double* V, const double* Q, double beta, float* table_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  int rc = build_table(ctx, V, Q, beta);
  if (rc) return rc;
  if (table_out) CK(cudaMemcpyAsync(table_out, ctx->d_table, 32 * (size_t)ctx->g.n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PIQL_OK;
}

int32_t piql_policy_table_async(piql_ctx* ctx, const double* V, const double* Q, double beta) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  NEED(V != nullptr, PIQL_ERR_ARG, "V is null");
  NEED(beta == beta, PIQL_ERR_ARG, "beta is NaN");
  const long long n = ctx->g.n;
  const double vg = V[ctx->g.goal];
  NEED(vg == vg, PIQL_ERR_ARG, "V(goal) is NaN");
  cudaStream_t st = ctx->stream;
  if (!ctx->ustream) {
    CK(cudaStreamCreateWithFlags(&ctx->ustream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_retire_cur, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_retire_spare, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_table_ready, cudaEventDisableTiming));
  }
  if (!ctx->d_table_spare) CK(cudaMalloc((void**)&ctx->d_table_spare, 32 * (size_t)n));
  if (ctx->d_table16 && !ctx->d_table16_spare) CK(cudaMalloc((void**)&ctx->d_table16_spare, 16 * (size_t)n));
  CK(grow(ctx->d_upV, ctx->up_cap, n));
  if (Q) CK(grow(ctx->d_upQ, ctx->upq_cap, 4 * n));
  cudaStream_t us = ctx->ustream;
  // the spare buffer was the current table until the previous call: every launch that may read it was enqueued before the
  // event recorded then; the upload staging is free once the previous build has run (stream order on the upload stream)
  if (ctx->spare_retired) CK(cudaStreamWaitEvent(us, ctx->ev_retire_spare, 0));
  CK(cudaMemcpyAsync(ctx->d_upV, V, 8 * (size_t)n, cudaMemcpyHostToDevice, us));
  if (Q) CK(cudaMemcpyAsync(ctx->d_upQ, Q, 32 * (size_t)n, cudaMemcpyHostToDevice, us));
  ctx->t16_valid_spare = ctx->d_table16_spare != nullptr && Q == nullptr;
  launch_policy_table(ctx->g, ctx->d_cell, ctx->d_upV, Q ? ctx->d_upQ : nullptr, beta, 1.0, ctx->d_table_spare, us, nullptr,
                      ctx->t16_valid_spare ? ctx->d_table16_spare : nullptr);
  CK_LAST();
  CK(cudaEventRecord(ctx->ev_table_ready, us));
  // retire the current table: whatever is enqueued on the context's stream so far may still read it
  CK(cudaEventRecord(ctx->ev_retire_cur, st));
  CK(cudaStreamWaitEvent(st, ctx->ev_table_ready, 0));
  std::swap(ctx->d_table, ctx->d_table_spare);
  std::swap(ctx->d_table16, ctx->d_table16_spare);
  std::swap(ctx->t16_valid, ctx->t16_valid_spare);
  std::swap(ctx->ev_retire_cur, ctx->ev_retire_spare);
  ctx->spare_retired = true;
  ctx->v_goal = vg;
  ctx->has_table = true;
  return PIQL_OK;
}

int32_t piql_policy_table16(piql_ctx* ctx, uint32_t* table16_out, int32_t* in_use_out) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(ctx->has_table, PIQL_ERR_STATE, "no policy table (call piql_policy_table first)");
  CK(cudaSetDevice(ctx->device));
  if (in_use_out) *in_use_out = ctx->t16_valid ? 1 : 0;
  if (table16_out && ctx->t16_valid)
    CK(cudaMemcpyAsync(table16_out, ctx->d_table16, 16 * (size_t)ctx->g.n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PIQL_OK;
}

int32_t piql_policy_table_set(piql_ctx* ctx, const float* table, double v_goal) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  NEED(table != nullptr, PIQL_ERR_ARG, "table is null");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->d_table, table, 32 * (size_t)ctx->g.n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->t16_valid = false;  // a caller's table need not have action-independent fitness entries
  ctx->v_goal = v_goal;
  ctx->has_table = true;
  return PIQL_OK;
}

int32_t piql_mc_logz(piql_ctx* ctx, const double* V, const double* Q, double beta, const piql_mc_opts* opts,
                     double* logz_sa_out, uint64_t* count_sa_out, double* logz_s_out, piql_mc_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  int rc = mc_check_opts(ctx, opts);
  if (rc) return rc;
  rc = build_table(ctx, V, Q, beta);
  if (rc) return rc;
  ctx->tev_used = 0;
  rc = mc_reset(ctx, opts);
  if (rc) return rc;
  rc = mc_rollout(ctx, opts, nullptr, 0);
  if (rc) return rc;
  return mc_finalize(ctx, logz_sa_out, count_sa_out, logz_s_out, stats);
}

int32_t piql_mc_update_v(piql_ctx* ctx, const double* V, double beta, const piql_mc_opts* opts, int32_t niter,
                         double* vnew_out, piql_mc_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  NEED(V != nullptr && vnew_out != nullptr, PIQL_ERR_ARG, "V / vnew_out is null");
  NEED(beta == beta && beta != 0.0, PIQL_ERR_ARG, "beta must be non-zero and not NaN");
  NEED(niter > 0, PIQL_ERR_ARG, "niter must be positive");
  int rc = mc_check_opts(ctx, opts);
  if (rc) return rc;
  NEED(opts->injected == nullptr, PIQL_ERR_ARG, "injected uniforms are for single rollouts, not for the update loop");
  rc = ensure_vi(ctx);
  if (rc) return rc;
  const int n = ctx->g.n;
  cudaStream_t st = ctx->stream;
  NEED(V[ctx->g.goal] == V[ctx->g.goal], PIQL_ERR_ARG, "V(goal) is NaN");
  CK(cudaMemcpyAsync(ctx->d_in0, V, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_convert_in(ctx->d_in0, ctx->d_in1, n, PIQL_F64, st);  // NaN at walls -> 0; V lives in d_in1 from here on
  CK_LAST();
  ctx->tev_used = 0;
  unsigned long long transitions = 0, paths = 0;
  piql_mc_opts o = *opts;
  for (int it = 0; it < niter; it++) {
    // the value at the goal enters the kernel as a scalar: fetch it (8 bytes; also the loop's only host sync)
    CK(cudaMemcpyAsync(ctx->h_scalars + 4, ctx->d_in1 + ctx->g.goal, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->v_goal = ctx->h_scalars[4];
    ctx->t16_valid = ctx->d_table16 != nullptr;
    launch_policy_table(ctx->g, ctx->d_cell, ctx->d_in1, nullptr, beta, 1.0, ctx->d_table, st, nullptr, ctx->d_table16);
    ctx->has_table = true;
    rc = mc_reset(ctx, &o);
    if (rc) return rc;
    rc = mc_rollout(ctx, &o, nullptr, 0);
    if (rc) return rc;
    launch_mc_apply_update(ctx->acc_max, ctx->acc_sum, ctx->acc_cnt, ctx->d_starts, ctx->nstarts, beta, ctx->d_in1, st);
    CK_LAST();
    CK(cudaMemcpyAsync(ctx->h_transitions, ctx->d_transitions, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    transitions += *ctx->h_transitions;
    paths += (unsigned long long)o.npaths;
    o.seed += 1;                    // fresh, reproducible stream per iteration
  }
  launch_convert_out(ctx->d_in1, ctx->d_cell, ctx->d_out, n, PIQL_F64, st);
  CK_LAST();
  CK(cudaMemcpyAsync(vnew_out, ctx->d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (stats) {
    double ro = 0.0, re = 0.0;
    rc = mc_timing(ctx, &ro, &re, nullptr, false);
    if (rc) return rc;
    stats->transitions = transitions; stats->paths = paths; stats->rollout_seconds = ro; stats->device_seconds = ro + re;
  }
  return PIQL_OK;
}

int32_t piql_trace_paths(piql_ctx* ctx, const piql_mc_opts* opts, int64_t max_steps, int32_t* triples_out,
                         float* logw_out, uint32_t* meta_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  int rc = mc_check_opts(ctx, opts);
  if (rc) return rc;
  NEED(max_steps > 0 && triples_out, PIQL_ERR_ARG, "max_steps must be positive and triples_out non-null");
  NEED(opts->npaths * max_steps < (1LL << 28), PIQL_ERR_ARG, "trace too large (npaths*max_steps must be < 2^28)");
  rc = mc_reset(ctx, opts);
  if (rc) return rc;
  int* d_trace = nullptr;
  const size_t nt = (size_t)opts->npaths * (size_t)max_steps * 3;
  CK(cudaMalloc((void**)&d_trace, 4 * (nt ? nt : 1)));
  cudaError_t e = cudaMemsetAsync(d_trace, 0xff, 4 * nt, ctx->stream);  // -1 padding
  rc = e == cudaSuccess ? mc_rollout(ctx, opts, d_trace, max_steps) : fail(ctx, PIQL_ERR_CUDA, cudaGetErrorString(e));
  if (rc == PIQL_OK) {
    e = cudaMemcpyAsync(triples_out, d_trace, 4 * nt, cudaMemcpyDeviceToHost, ctx->stream);
    if (logw_out || meta_out) {
      rc = ensure_split(ctx, opts->npaths ? opts->npaths : 1);
      if (rc == PIQL_OK) launch_split_records(ctx->d_rec, opts->npaths, ctx->d_logw, ctx->d_meta, ctx->stream);
    }
    if (rc == PIQL_OK && e == cudaSuccess && logw_out) e = cudaMemcpyAsync(logw_out, ctx->d_logw, 4 * (size_t)opts->npaths, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == PIQL_OK && e == cudaSuccess && meta_out) e = cudaMemcpyAsync(meta_out, ctx->d_meta, 4 * (size_t)opts->npaths, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, PIQL_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(d_trace);
  return rc;
}

int32_t piql_path_estimates(piql_ctx* ctx, const double* V, const double* Q, const piql_mc_opts* opts, int32_t nsteps,
                            float* fitness_out, float* direct_out, int32_t* triples_out, float* table_out, float* kl_out) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded (call piql_grid_load first)");
  int rc = mc_check_opts(ctx, opts);
  if (rc) return rc;
  NEED(nsteps > 0, PIQL_ERR_ARG, "nsteps must be positive");
  NEED(fitness_out && direct_out, PIQL_ERR_ARG, "fitness_out / direct_out is null");
  NEED(opts->npaths * (long long)nsteps < (1LL << 28), PIQL_ERR_ARG, "too many running totals (npaths*nsteps must be < 2^28)");
  NEED(!opts->injected || opts->injected_steps >= nsteps, PIQL_ERR_ARG, "injected_steps must cover nsteps");
  NEED(V != nullptr, PIQL_ERR_ARG, "V is null");
  rc = mc_setup_slots(ctx, opts);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const int n = ctx->g.n;
  const size_t nt = (size_t)opts->npaths * (size_t)nsteps;
  float *d_kl = nullptr, *d_fit = nullptr;
  int* d_tr = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_kl, 4 * (size_t)n);
  if (e == cudaSuccess) {
    // beta = 1: the record's weight column is fitness(s,a) itself; kl_cost per cell beside it
    rc = build_table(ctx, V, Q, 1.0, d_kl);
    ctx->has_table = false;  // not the sampler's table of whatever beta the caller asked for earlier
    if (rc) { cudaFree(d_kl); return rc; }
  }
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_fit, 8 * (nt ? nt : 1));
  if (e == cudaSuccess && triples_out) e = cudaMalloc((void**)&d_tr, 12 * (nt ? nt : 1));
  const float* inj_dev = nullptr;
  if (e == cudaSuccess && opts->injected) {
    const size_t nf = (size_t)opts->npaths * (size_t)opts->injected_steps * 4;
    e = grow(ctx->d_inj, ctx->inj_cap, nf);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_inj, opts->injected, 4 * nf, cudaMemcpyHostToDevice, st);
    inj_dev = ctx->d_inj;
  }
  if (e == cudaSuccess) {
    PathEstArgs a{};  // d_in0 still holds the caller's V in double
    a.table = ctx->d_table; a.kl = d_kl; a.V = ctx->d_in0;
    a.starts = ctx->d_starts; a.nstarts = ctx->nstarts; a.path0 = opts->path0; a.npaths = opts->npaths;
    a.nsteps = nsteps; a.H = ctx->g.H; a.goal = ctx->g.goal;
    a.rho = ctx->g.randomness; a.rho_k = (uint32_t)std::ceil(a.rho * 16777216.0);
    a.rwd_goal = (float)ctx->g.reward; a.rwd_step = (float)(-ctx->g.step_cost);
    a.keys = philox_expand((uint32_t)opts->seed, (uint32_t)(opts->seed >> 32));
    a.injected = inj_dev; a.injected_steps = opts->injected_steps;
    a.fit = d_fit; a.dir = d_fit + nt; a.trace = d_tr;
    launch_path_estimates(a, ctx->num_sms, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(fitness_out, d_fit, 4 * nt, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(direct_out, d_fit + nt, 4 * nt, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && triples_out) e = cudaMemcpyAsync(triples_out, d_tr, 12 * nt, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && table_out) e = cudaMemcpyAsync(table_out, ctx->d_table, 32 * (size_t)n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && kl_out) e = cudaMemcpyAsync(kl_out, d_kl, 4 * (size_t)n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  for (void* p : {(void*)d_kl, (void*)d_fit, (void*)d_tr}) if (p) cudaFree(p);
  if (e != cudaSuccess) return fail(ctx, PIQL_ERR_CUDA, std::string("piql_path_estimates: ") + cudaGetErrorString(e));
  return PIQL_OK;
}

int32_t piql_mc_reset_dev(piql_ctx* ctx, const piql_mc_opts* opts) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  int rc = mc_check_opts(ctx, opts);
  if (rc) return rc;
  return mc_reset(ctx, opts);
}
int32_t piql_mc_rollout_dev(piql_ctx* ctx, const piql_mc_opts* opts) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  NEED(ctx->nstarts > 0, PIQL_ERR_STATE, "call piql_mc_reset_dev first");
  int rc = mc_check_opts(ctx, opts);
  if (rc) return rc;
  return mc_rollout(ctx, opts, nullptr, 0);
}
int32_t piql_mc_stats_dev(piql_ctx* ctx, void** lse_max_dev, void** lse_sum_dev, void** count_dev, int64_t* n_out) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(ctx->nstarts > 0 && ctx->acc_max, PIQL_ERR_STATE, "call piql_mc_reset_dev first");
  if (lse_max_dev) *lse_max_dev = ctx->acc_max;
  if (lse_sum_dev) *lse_sum_dev = ctx->acc_sum;
  if (count_dev) *count_dev = ctx->acc_cnt;
  if (n_out) *n_out = ctx->nstarts * NA;
  return PIQL_OK;
}
int32_t piql_mc_rescale_dev(piql_ctx* ctx, const void* global_max_dev) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  NEED(ctx->nstarts > 0 && ctx->acc_max, PIQL_ERR_STATE, "call piql_mc_reset_dev first");
  NEED(global_max_dev != nullptr, PIQL_ERR_ARG, "global_max_dev is null");
  CK(cudaSetDevice(ctx->device));
  launch_mc_rescale(ctx->acc_max, ctx->acc_sum, (const float*)global_max_dev, ctx->nstarts * NA, ctx->stream);
  CK_LAST();
  return PIQL_OK;
}
int32_t piql_peer_export(piql_ctx* ctx, piql_peer_handles* out) {
  if (!ctx || !out) return PIQL_ERR_ARG;
  NEED(ctx->nstarts > 0 && ctx->acc_max, PIQL_ERR_STATE, "call piql_mc_reset_dev first");
  CK(cudaSetDevice(ctx->device));
  PeerState& ps = ctx->peer;
  peer_close(ctx);
  if (!ps.own_flags) {
    CK(cudaMalloc((void**)&ps.own_flags, sizeof(unsigned long long) * 2 * PEER_MAX));
    CK(cudaMalloc((void**)&ps.done, sizeof(unsigned int)));
    CK(cudaMallocHost((void**)&ps.err, sizeof(int)));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemset(ps.own_flags, 0, sizeof(unsigned long long) * 2 * PEER_MAX));
  CK(cudaMemset(ps.done, 0, sizeof(unsigned int)));
  *ps.err = 0;
  ps.epoch = 0;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle blob size");
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->lse_max, ctx->acc_max));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->lse_sum, ctx->acc_sum));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->count, ctx->acc_cnt));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->flags, ps.own_flags));
  out->n = ctx->nstarts * NA;
  ps.bound_mx = ctx->acc_max;
  return PIQL_OK;
}
int32_t piql_peer_import(piql_ctx* ctx, int32_t rank, int32_t world, const piql_peer_handles* all) {
  if (!ctx || !all) return PIQL_ERR_ARG;
  NEED(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, PIQL_ERR_ARG, "rank / world out of range (at most 8 GPUs)");
  PeerState& ps = ctx->peer;
  NEED(ps.own_flags && ps.bound_mx == ctx->acc_max, PIQL_ERR_STATE, "call piql_peer_export first");
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < world; r++)
    NEED(all[r].n == ctx->nstarts * NA, PIQL_ERR_ARG, "ranks disagree on the accumulator length (different start lists)");
  ps.rank = rank; ps.world = world;
  for (int r = 0; r < world; r++) {
    if (r == rank) {
      ps.mx[r] = ctx->acc_max; ps.sm[r] = ctx->acc_sum; ps.cn[r] = ctx->acc_cnt; ps.flags[r] = ps.own_flags;
      continue;
    }
    void* p[4] = {nullptr, nullptr, nullptr, nullptr};
    const uint8_t* h[4] = {all[r].lse_max, all[r].lse_sum, all[r].count, all[r].flags};
    for (int q = 0; q < 4; q++) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, h[q], sizeof(hd));
      cudaError_t e = cudaIpcOpenMemHandle(&p[q], hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        peer_close(ctx);
        return fail(ctx, PIQL_ERR_CUDA, std::string("cudaIpcOpenMemHandle (is the peer GPU NVLink / P2P reachable?): ") +
                                            cudaGetErrorString(e));
      }
      ps.opened[ps.nopened++] = p[q];
    }
    ps.mx[r] = (float*)p[0]; ps.sm[r] = (float*)p[1]; ps.cn[r] = (uint32_t*)p[2]; ps.flags[r] = (unsigned long long*)p[3];
  }
  ps.valid = true;
  return PIQL_OK;
}
int32_t piql_mc_allreduce_peer(piql_ctx* ctx) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  PeerState& ps = ctx->peer;
  NEED(ps.valid && ps.bound_mx == ctx->acc_max, PIQL_ERR_STATE, "peer mappings are missing or stale (piql_peer_export / piql_peer_import)");
  NEED(*ps.err == 0, PIQL_ERR_CUDA, "an earlier peer exchange timed out waiting for another rank");
  CK(cudaSetDevice(ctx->device));
  ps.epoch++;
  launch_lse_allreduce_peer(ps, ctx->nstarts, ctx->num_sms, ctx->stream);
  CK_LAST();
  return PIQL_OK;
}
int32_t piql_peer_close(piql_ctx* ctx) {
  if (!ctx) return PIQL_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  peer_close(ctx);
  return PIQL_OK;
}
int32_t piql_mc_finalize(piql_ctx* ctx, double* logz_sa_out, uint64_t* count_sa_out, double* logz_s_out,
                         piql_mc_stats* stats) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  return mc_finalize(ctx, logz_sa_out, count_sa_out, logz_s_out, stats);
}
int32_t piql_mc_finalize_async(piql_ctx* ctx, void* logz_sa_out, void* count_sa_out, void* logz_s_out, int32_t f32_outputs) {
  if (!ctx) return PIQL_ERR_ARG;
  LaunchScope ls(ctx);
  CK(cudaSetDevice(ctx->device));
  NEED(ctx->has_grid, PIQL_ERR_STATE, "no grid loaded");
  return mc_finalize_async(ctx, logz_sa_out, count_sa_out, logz_s_out, f32_outputs != 0);
}
int32_t piql_mc_finalize_wait(piql_ctx* ctx, piql_mc_stats* stats_out) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(ctx->fin_pending, PIQL_ERR_STATE, "no asynchronous finalisation in flight (call piql_mc_finalize_async first)");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev_done));
  ctx->fin_pending = false;
  if (ctx->peer.err && *ctx->peer.err)
    return fail(ctx, PIQL_ERR_CUDA, "the peer exchange timed out waiting for another rank (did every rank call it?)");
  if (stats_out) {
    stats_out->transitions = *ctx->h_fin_transitions;
    stats_out->paths = ctx->fin_paths;
    stats_out->rollout_seconds = 0.0;   // the per-launch event timers belong to the stream that may still be running
    stats_out->device_seconds = 0.0;
  }
  return PIQL_OK;
}
int32_t piql_mc_timing(piql_ctx* ctx, double* rollout_seconds, double* reduce_seconds, int64_t* launches) {
  if (!ctx) return PIQL_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return mc_timing(ctx, rollout_seconds, reduce_seconds, launches, true);
}
int32_t piql_mc_transitions_total(piql_ctx* ctx, uint64_t* total_out) {
  if (!ctx || !total_out) return PIQL_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->h_transitions, ctx->d_transitions, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *total_out = ctx->h_transitions[0] + ctx->h_transitions[1];
  return PIQL_OK;
}
int32_t piql_mc_records(piql_ctx* ctx, float* logw_out, uint32_t* meta_out, int64_t n) {
  if (!ctx) return PIQL_ERR_ARG;
  NEED(n >= 0 && n <= (ctx->records_split ? ctx->split_cap : ctx->rec_cap), PIQL_ERR_ARG,
       "more records requested than the last rollout produced");
  CK(cudaSetDevice(ctx->device));
  LaunchScope ls(ctx);
  if (!ctx->records_split) {  // the grid sampler keeps {log-weight, meta} interleaved: split them for the host
    int rc = ensure_split(ctx, n ? n : 1);
    if (rc) return rc;
    launch_split_records(ctx->d_rec, n, ctx->d_logw, ctx->d_meta, ctx->stream);
    CK_LAST();
  }
  if (logw_out) CK(cudaMemcpyAsync(logw_out, ctx->d_logw, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (meta_out) CK(cudaMemcpyAsync(meta_out, ctx->d_meta, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PIQL_OK;
}

}  // extern "C"
