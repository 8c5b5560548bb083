// internal.h -- host-side interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace piql {

// KIND_Z: the reference's linear-space twin of KIND_LOGZ (update_z!, value_iteration.jl:23-40); swept by the plain
// per-cell kernels only (resident solver or one launch per sweep), never by the tiled kernels.
enum Kind { KIND_NU = 0, KIND_LOGZ = 1, KIND_GAP = 2, KIND_Z = 3 };

// Device state of one soft-VI solve.  All value arrays hold n = H*W elements of the sweep
// precision (float or double), column-major like the wall array.
struct ViState {
  int precision;      // PIQL_F64 / PIQL_F32
  int kind;
  double beta, reward_scale;
  double v_goal;      // V(goal): logz at the terminal cell is 0 - V(goal)
  void* x[2];         // Jacobi ping-pong iterate; x[cur] is current
  int cur;
  void* V;            // fixed V (logz, gap)
  void* aux4;         // logz: A[4] per cell; gap: pi[4] per cell
  void* aux1;         // gap: D per cell
  double* partial;    // per-block residual partials
  double* residual;   // device scalar
  int nblocks;
  bool fast;          // fp32 only: MUFU ex2/lg2 approximations in the tiled sweeps (PIQL_F32_FAST)
};

// tile.cu: temporally blocked sweeps on a padded copy of the iterate ---------------------------
constexpr int TILE_KH_MAX = 8;  // deepest temporal blocking (sweeps per round = ghost depth of the padded layout): 4 or 8 per plan
struct TilePlan {
  int TJ, EJ;            // interior / staged columns per tile (EJ = TJ + 2*kh); rows are 128 - 2*kh / 128
  int kh;                // Jacobi sweeps per round = halo width = ghost frame of the padded layout (4 or 8)
  int planes;            // constant planes staged in shared memory per round (fast fp32 log Z: 4, gap: 5), 0: read from global
  int tiles_i, tiles_j;
  int nw;                // warps per CTA
  int ld, Wp;            // padded leading dimension and column count
  long long np;          // ld * Wp
  size_t smem;           // dynamic shared memory per CTA
  int ctas_per_sm;
  int gi;                // ghost rows of the padded layout (= kh for the tiled plans)
  // column-resident schedule (col_resident_kernel: one CTA per SM keeps a slab of whole columns in shared memory for a whole
  // batch of sweeps and exchanges one column per sweep with each neighbour): geometry in padded coordinates
  int colres;            // 1: this plan
  int r0, NS;            // first resident row (multiple of 4), 128-row strips per column
  int c0, ncta;          // first column of CTA 0, CTAs
  int wbase, wrem, wmax; // CTA c owns wbase + (c < wrem) columns; wmax = the widest slab
  int nwc, G;            // compute warps (nw less the mail warps), groups per warp of the work list
};
// Halo exchange of the multi-GPU soft VI, carried by the tile kernel itself: the CTAs of the slab's first / last tile column
// wait for the neighbour's counter before they stage their tile, and after the write-back also store their edge columns
// into the neighbour's array and (the last of them) raise the neighbour's counter.  All null / 0 on one GPU.
struct TileHalo {
  const unsigned* in_left = nullptr;   // local counters the neighbours raise
  const unsigned* in_right = nullptr;
  unsigned wait_epoch = 0;             // 0: nothing to wait for (first round after packing)
  void* out_left = nullptr;            // the neighbours' copies of the buffer this launch writes
  void* out_right = nullptr;
  unsigned* flag_left = nullptr;       // counters on the neighbours
  unsigned* flag_right = nullptr;
  unsigned* done = nullptr;            // local: edge CTAs that have pushed, [0] left, [1] right
  unsigned post_epoch = 0;
  int tj_first = -1, tj_last = -1;
  int* err = nullptr;
  unsigned long long* wait_stats = nullptr;  // [0] nanoseconds edge CTAs spent waiting for a neighbour's counter, [1] waits
};
struct TileBuffers {
  void* x[2] = {nullptr, nullptr};  // padded ping-pong iterate
  void* aux = nullptr;              // five padded planes (logz: exp(A_a - max A), max A; gap: pi_a, D)
  uint8_t* mask = nullptr;          // padded cell bytes (ghost frame = wall)
  double* partial = nullptr;        // per-tile residual partials
  size_t cap_x = 0, cap_aux = 0, cap_mask = 0;
  long long cap_partial = 0;
  TilePlan plan{};
  int cur = 0;
  bool active = false;
  // CUtensorMap descriptors (opaque here) of x[0], x[1] and mask: boxes of 128 x EJ elements for the TMA staging
  alignas(64) unsigned char tmx[2][128] = {};
  alignas(64) unsigned char tmm[128] = {};
  alignas(64) unsigned char tmp[128] = {};  // 3-D map over the constant planes (fast fp32 log Z / gap: planes staged per round)
  bool use_tma = false;
  // persistent variant (all tiles resident at once): one cooperative launch per batch of sweeps
  bool persistent = false;
  unsigned* flags = nullptr;   // rounds completed per tile
  long long cap_flags = 0;
  int* err = nullptr;          // pinned host word: set when a neighbour wait timed out (copied from d_err after a launch)
  int* d_err = nullptr;        // device word the kernel polls and sets
  // column-resident schedule: L2 mailboxes of the per-sweep column exchange ({word, sequence} pairs) and the sequence counter
  unsigned long long* mail = nullptr;
  size_t cap_mail = 0;
  unsigned col_seq = 0;
  void* sched = nullptr;       // work list of a sweep (uint2 [2 shapes][nwc][G]), built by tile_setup
  size_t cap_sched = 0;
  // multi-GPU column slabs: this rank sweeps tile columns [tj0, tj1) (piql_vi_shard_*)
  bool sharded = false;
  int tj0 = 0, tj1 = 0;
  TileHalo halo;               // set per launch by tile_sweeps_sharded
  // shapes found fastest on the device, one per problem key (grid, precision, kind): a context that alternates between
  // nu / log Z / gap solves or precisions on one grid times each of them once
  static constexpr int TUNED_MAX = 24;
  int ntuned = 0, tuned_next = 0;
  long long tuned_key[TUNED_MAX] = {};
  TilePlan tuned_plan[TUNED_MAX] = {};
  bool tuned_persistent[TUNED_MAX] = {};
};
TilePlan make_tile_plan(int H, int W, int precision, int kind, bool fast, int num_sms);
TilePlan choose_tile_plan(int H, int W, int precision, int kind, bool fast, int num_sms, bool* persistent);
cudaError_t tile_begin(const GridDev& g, const uint8_t* cell, const ViState& vs, TileBuffers& tb, int num_sms, cudaStream_t st);
cudaError_t tile_sweeps(const GridDev& g, const ViState& vs, TileBuffers& tb, long long nsweeps, bool want_res, cudaStream_t st);
void tile_unpack(const GridDev& g, const ViState& vs, const TileBuffers& tb, cudaStream_t st);
void tile_free(TileBuffers& tb);

constexpr int SWEEP_THREADS = 256;
constexpr int RESIDENT_MAX_CELLS = 8192;  // single-CTA resident solver limit

// backup.cu ---------------------------------------------------------------------------------
void launch_cellinfo(const uint8_t* walls, const GridDev& g, uint8_t* cell, cudaStream_t st);
// Converts double input to the sweep precision (or copies).
void launch_convert_in(const double* src, void* dst, int n, int precision, cudaStream_t st);
// Writes double output with NaN at walls.
void launch_convert_out(const void* src, const uint8_t* cell, double* dst, int n, int precision, cudaStream_t st);
void launch_fill(void* dst, double value, int n, int precision, cudaStream_t st);
// Precompute the per-cell constants of the logz / gap sweeps from V (and Q, nu).
void launch_prep(const GridDev& g, const uint8_t* cell, ViState& vs, const double* V, const double* Q, const double* nu,
                 cudaStream_t st);
// One Jacobi sweep x[cur] -> x[cur^1]; with_residual: also per-block partials + final sum.
void launch_sweep(const GridDev& g, const uint8_t* cell, ViState& vs, bool with_residual, cudaStream_t st);
// Whole solve in one CTA (n <= RESIDENT_MAX_CELLS): returns through dev_out {sweeps, residual}.
void launch_resident_solve(const GridDev& g, const uint8_t* cell, ViState& vs, double tol, long long max_sweeps,
                           double* dev_out /*2 doubles*/, cudaStream_t st);
// linear: logz holds Z (z_updateV, value_iteration.jl:53-59) instead of log Z
void launch_update_v(const double* V, const void* logz, const uint8_t* cell, double beta, double* out, int n,
                     int precision, cudaStream_t st, bool linear = false);
size_t resident_smem_bytes(int n, int precision);
// One CTA per problem (beta chain of logz_updateV solves, or single nu / gap solves), everything in shared memory.
size_t resident_batch_smem_bytes(int n, int precision);
void launch_resident_batch(const GridDev& g, const uint8_t* cell, int kind, int precision, double reward_scale, int nproblems,
                           const double* Vin, long long vin_stride, const double* nu, const double* betas, int nupdates,
                           double tol, long long max_sweeps, double* out, double* stats, cudaStream_t st);
int configure_backup_kernels();
void launch_residual_final(const double* partial, int nblocks, double* out, cudaStream_t st);

// rollout.cu --------------------------------------------------------------------------------
struct RolloutArgs {
  const float4* table;   // 2 x float4 per cell
  const uint4* table16;  // nullptr, or the 16-byte records (V-determined Q on grids beyond L2): the production kernel reads these
  long long table_bytes; // 32 * cells
  int goal;
  int H;
  double gamma, rho;        // the problem's Float64 gamma and randomness: injected uniforms are compared with them as doubles
  float vgoal, lgam, l1mg;
  uint32_t gamma_k, rho_k;  // ceil(gamma * 2^24), ceil(randomness * 2^24): the same tests on the 24-bit integers behind the Philox uniforms
  int estimator;
  long long horizon;
  long long path0, npaths;
  const int* starts;
  long long nstarts;
  PhiloxKeys keys;       // round keys of Philox4x32-10 for (seed lo, seed hi)
  const float* injected;
  long long injected_steps;
  uint2* rec;            // [npaths] per-path records {log-weight bits, a0 | len << 2}: one 8-byte store per finished path
  int* trace;            // optional [npaths][trace_steps][3]
  long long trace_steps;
};
// kl_out (optional): kl_cost(ctrl, s, Q) per cell, path_functions.jl:17-20
void launch_policy_table(const GridDev& g, const uint8_t* cell, const double* V, const double* Q, double beta,
                         double reward_scale, float4* table, cudaStream_t st, float* kl_out = nullptr, uint4* table16 = nullptr);
// Grids from this many cells on get the 16-byte records as well (their 32-byte table, >= 128 MB, does not stay in L2);
// PIQL_ROLL_REC16 = 0 / 1 overrides (tests, A/B timing).
constexpr long long ROLL_REC16_MIN_CELLS = 4LL << 20;
struct PathEstArgs {
  const float4* table;   // policy records built with beta = 1 (bf = fitness)
  const float* kl;       // kl_cost per cell
  const double* V;       // the caller's V (double, device): V(s0) opens both totals
  const int* starts;
  long long nstarts, path0, npaths;
  int nsteps, H, goal;
  double rho;
  float rwd_goal, rwd_step;
  uint32_t rho_k;
  PhiloxKeys keys;
  const float* injected;
  long long injected_steps;
  float* fit;            // [npaths][nsteps]
  float* dir;            // [npaths][nsteps]
  int* trace;            // optional [npaths][nsteps][3]
};
void launch_path_estimates(const PathEstArgs& a, int num_sms, cudaStream_t st);
void launch_rollout(const RolloutArgs& a, int num_sms, cudaStream_t st);
// Reduce records into (max, sum, count)[nstarts*4], merged into the accumulators; adds the
// transition count to *transitions.
void launch_reduce(const uint2* rec, long long path0, long long npaths, long long nstarts,
                   float* acc_max, float* acc_sum, uint32_t* acc_cnt, float* scratch, size_t scratch_floats,
                   unsigned long long* transitions, int num_sms, cudaStream_t st);
size_t reduce_scratch_floats(long long npaths, long long nstarts);
// De-interleaves the first n records into separate log-weight / meta arrays (piql_mc_records, piql_trace_paths).
void launch_split_records(const uint2* rec, long long n, float* logw, uint32_t* meta, cudaStream_t st);
void launch_mc_reset(float* acc_max, float* acc_sum, uint32_t* acc_cnt, long long n, cudaStream_t st);
void launch_fold_counter(unsigned long long* counters, cudaStream_t st);
void launch_mc_rescale(float* acc_max, float* acc_sum, const float* gmax, long long n, cudaStream_t st);
// Scatter accumulators into per-cell outputs (NaN / 0 elsewhere already set by the caller).
void launch_mc_finalize(const float* acc_max, const float* acc_sum, const uint32_t* acc_cnt, const int* starts,
                        long long nstarts, double* logz_sa, unsigned long long* count_sa, double* logz_s,
                        cudaStream_t st);
void launch_mc_finalize_f32(const float* acc_max, const float* acc_sum, const uint32_t* acc_cnt, const int* starts,
                            long long nstarts, float* logz_sa, uint32_t* count_sa, float* logz_s, cudaStream_t st);
void launch_fill_f64(double* dst, double v, long long n, cudaStream_t st);
void launch_fill_f32(float* dst, float v, long long n, cudaStream_t st);
void launch_mc_apply_update(const float* acc_max, const float* acc_sum, const uint32_t* acc_cnt, const int* starts,
                            long long nstarts, double beta, double* V, cudaStream_t st);

// table.cu: soft VI on a tabulated ControlProblem (CSR over (state, action)) -------------------------
struct TableDev {
  long long ns;
  int na;
  const long long* ptr;   // [ns*na + 1]
  const int* succ;        // [nnz]
  const double* prob;     // [nnz]
  const double* rew;      // [nnz] reward_function(s, a, s')
  const double* lprior;   // [ns*na] log action_prior(s, a)
  const uint8_t* term;    // [ns]
  double gamma, lgamma, l1mgamma;
};
void launch_table_prep(const TableDev& t, int kind, double scale, double beta, const double* V, const double* nu, double* aux,
                       double* aux1, cudaStream_t st, const double* Q = nullptr);
// One Jacobi sweep X -> Y; partial != nullptr: also the residual sum |new - old| into *residual.
void launch_table_sweep(const TableDev& t, int kind, double scale, const double* X, double* Y, const double* V, const double* aux,
                        const double* aux1, double* partial, double* residual, cudaStream_t st);
void launch_table_update_v(long long ns, const double* V, const double* logz, double beta, double* out, cudaStream_t st);
int table_blocks(long long ns);
// Monte-Carlo sampling on the table: sampler tables from V, rollout (estimator A), reduction + finalisation
void launch_table_policy(const TableDev& t, double scale, double beta, const double* V, const double* Q, float* cdf, float* bf,
                         float* cum, float* vterm, cudaStream_t st);
void launch_table_rollout(const TableDev& t, const float* cdf, const float* bf, const float* cum, const float* vterm, double gamma,
                          long long horizon, long long path0, long long npaths, const int* starts, long long nstarts,
                          uint64_t seed, float* logw, uint32_t* meta, int num_sms, cudaStream_t st);
void launch_table_reduce(const float* logw, const uint32_t* meta, long long path0, long long npaths, const int* starts,
                         long long nstarts, int na, int* omax, double* sum, unsigned long long* cnt,
                         unsigned long long* transitions, double* logz_sa, unsigned long long* count_sa, double* logz_s,
                         int num_sms, cudaStream_t st);

// peer.cu: the log-sum-exp all-reduce of the accumulators over NVLink peer memory -----------------
constexpr int PEER_MAX = 8;  // GPUs of one box
struct PeerState {
  int rank = 0, world = 0;
  float* mx[PEER_MAX] = {};
  float* sm[PEER_MAX] = {};
  uint32_t* cn[PEER_MAX] = {};
  unsigned long long* flags[PEER_MAX] = {};  // flags[r] on rank r: [2 phases][PEER_MAX source ranks]
  unsigned long long* own_flags = nullptr;
  unsigned int* done = nullptr;
  int* err = nullptr;                        // pinned host word the kernel writes on a timed-out wait
  unsigned long long epoch = 0;
  bool valid = false;
  const float* bound_mx = nullptr;           // accumulator allocation the handles were exported for
  void* opened[4 * PEER_MAX] = {};
  int nopened = 0;
};
void launch_lse_allreduce_peer(const PeerState& ps, long long nslots, int num_sms, cudaStream_t st);

// multi-GPU soft VI (tile.cu): peer mappings of the other ranks' padded ping-pong arrays and halo counters
struct ViPeer {
  int rank = 0, world = 0;
  void* x[2][PEER_MAX] = {};           // x[buffer][rank]
  unsigned* flags[PEER_MAX] = {};      // flags[r] on rank r: [0] rounds delivered by its left neighbour, [1] by its right one
  unsigned* own_flags = nullptr;       // [0], [1]: counters the neighbours raise; [2]: blocks of my push kernel that are done
  unsigned* done = nullptr;
  unsigned long long* wait_stats = nullptr;  // device: [0] ns waited for halos, [1] number of waits (piql_vi_shard_wait_stats)
  long long round = 0;                 // rounds completed since the iterate was (re)packed
  bool valid = false;
  void* opened[3 * PEER_MAX] = {};
  int nopened = 0;
};
cudaError_t tile_sweeps_sharded(const GridDev& g, const ViState& vs, TileBuffers& tb, ViPeer& vp, long long nsweeps, bool want_res,
                                cudaStream_t st);

// counts kernel launches made through the launch_* wrappers
extern thread_local long long g_launches;

}  // namespace piql
