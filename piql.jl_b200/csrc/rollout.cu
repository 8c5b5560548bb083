// rollout.cu -- Monte-Carlo path sampling under the soft policy pi_Q, with per-path log-weight
// accumulation and the log-sum-exp reduction into per-(start cell, first action) statistics.
//
// Replaces choose_action / path_transitions / fitness_reward_estimate (src/path_functions.jl:12-41),
// the propagator closure (src/Problems/GridWorld.jl:113-115) and StatsBase's sample(a, Weights).
// The estimator is the Monte-Carlo form of z_recursion (src/PIQL.jl:109-115): with
// pi_0 exp(beta cV + (1-beta) cQ) = pi_Q exp(beta fitness),
//     Z(s) = E_{a~pi_Q}[ exp(beta fitness(s,a)) (gamma E_{s'} Z(s') + 1 - gamma) ],  Z(goal) = exp(-V(goal)).
//
// Data layout: one 32-byte record per cell {c0,c1,c2,mask | bf0..bf3} = exactly one DRAM/L2
// sector per sampled transition; one 8-byte record per finished path {logw, a0|len<<2}, written with one 64-bit store.
// One trajectory per thread; a warp owns a contiguous range of path ids and refills finished
// lanes warp-synchronously (ballot + popc), so path -> lane assignment is deterministic.
#include <cstdlib>

#include "internal.h"

namespace piql {

// ---- policy table ----------------------------------------------------------------------------------
// Computed in double from the caller's V (and Q) and rounded once to float, like the oracle.
__global__ void __launch_bounds__(256)
policy_table_kernel(GridDev g, double reward_scale, double beta, const uint8_t* __restrict__ cell,
                    const double* __restrict__ V, const double* __restrict__ Q, float4* __restrict__ table,
                    float* __restrict__ kl_out, uint4* __restrict__ table16) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.n) return;
  const unsigned ci = cell[s];
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
  float kl = 0.f;
  if (!(ci & CELL_WALL)) {
    const double w_on = 1.0 * (1.0 - g.randomness) + g.randomness / NA;
    const double w_off = 0.0 * (1.0 - g.randomness) + g.randomness / NA;
    const double c = V[s];
    double nb[NA], rw[NA];
    const int off[NA] = {1, -1, g.H, -g.H};
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const bool blocked = (ci >> a) & 1u;
      nb[a] = blocked ? c : V[s + off[a]];
      // reward_function(s, a, s') looks at s' = take_step(s, a) alone (GridWorld.jl:116): a blocked move out of the goal
      // cell stays in the goal and earns the reward (only the walk of path_transitions ever stands on the goal)
      rw[a] = reward_scale * (((blocked ? s : s + off[a]) == g.goal) ? g.reward : -g.step_cost);
    }
    double cv[NA], q[NA], x[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
      double pv = 0.0, pr = 0.0;
#pragma unroll
      for (int aa = 0; aa < NA; aa++) {
        const double w = (a == aa) ? w_on : w_off;
        pv += w * nb[aa];
        pr += w * rw[aa];
      }
      cv[a] = g.gamma * pv - c + pr;                       // controlV   PIQL.jl:63-66
      q[a] = Q ? Q[(size_t)s * NA + a] : cv[a];
      x[a] = log(1.0 / NA) + q[a];
    }
    const double norm = lse4(x);                            // Qnormalization   PIQL.jl:72-75
    double w[NA], sum = 0.0, cw = 0.0;
#pragma unroll
    for (int a = 0; a < NA; a++) { w[a] = 0.25 * exp(q[a] - norm); sum += w[a]; }  // action_probability :83-85
    float cdf[3], bf[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
      cw += w[a];
      if (a < 3) cdf[a] = (float)(cw / sum);
      bf[a] = (float)(beta * (cv[a] - (q[a] - norm)));      // beta * fitness   PIQL.jl:87-89
    }
    r0 = make_float4(cdf[0], cdf[1], cdf[2], __uint_as_float(ci & CELL_BLOCK_MASK));
    r1 = make_float4(bf[0], bf[1], bf[2], bf[3]);
    if (kl_out) {  // kl_cost(ctrl, s, Q) = sum_a xlogx(p_a) - xlogy(p_a, prior)   path_functions.jl:17-20
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < NA; a++) acc += (w[a] == 0.0 ? 0.0 : w[a] * log(w[a])) - (w[a] == 0.0 ? 0.0 : w[a] * log(1.0 / NA));
      kl = (float)acc;
    }
  }
  table[2 * (size_t)s] = r0;
  table[2 * (size_t)s + 1] = r1;
  if (kl_out) kl_out[s] = kl;
  // 16-byte record (large grids, V-determined Q): with Q = controlV the fitness is -(Q - Qnormalization) + controlV = the
  // normalisation, the same for every action (PIQL.jl:87-89), so one float carries it (bf0, the 32-byte record's first
  // entry); the running sums are probabilities, < 2, so bits 30 and 31 of their floats are free and carry the four
  // blocked-move bits: {c0 | m01 << 30, c1 | m23 << 30, c2, bf0}.  Decoding gives back exactly c0, c1, c2, mask, bf0.
  if (table16) {
    const unsigned m = __float_as_uint(r0.w);
    table16[s] = make_uint4(__float_as_uint(r0.x) | ((m & 3u) << 30), __float_as_uint(r0.y) | (((m >> 2) & 3u) << 30),
                            __float_as_uint(r0.z), __float_as_uint(r1.x));
  }
}

void launch_policy_table(const GridDev& g, const uint8_t* cell, const double* V, const double* Q, double beta,
                         double reward_scale, float4* table, cudaStream_t st, float* kl_out, uint4* table16) {
  policy_table_kernel<<<(g.n + 255) / 256, 256, 0, st>>>(g, reward_scale, beta, cell, V, Q, table, kl_out, table16);
  g_launches++;
}

// ---- the reference's own path statistics ---------------------------------------------------------------
// fitness_reward_estimate / direct_reward_estimate (src/path_functions.jl:34-50) over path_transitions (:23-32): exactly
// nsteps sampled transitions per path -- no stop rule and no terminal handling, the walk goes on through the goal as the
// reference's does -- and after every transition the running totals
//   fit[t] = V(s0) + sum_{tau<=t} fitness(s_tau, a_tau)
//   dir[t] = V(s0) + sum_{tau<=t} [reward_function(s_tau, a_tau, s'_tau) - kl_cost(s_tau)]       (sampled reward)
// One path per thread (all paths have the same length, so there is nothing to refill); the per-step record is the
// sampler's 32-byte record built with beta = 1 plus one float of kl_cost per cell; fp32 running sums in the order
// written here, which the oracle repeats.  Uniform contract as the estimator's (word 1, the stop test, is drawn and unused).
template <bool INJ>
__global__ void __launch_bounds__(256)
path_estimates_kernel(const PathEstArgs p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < p.npaths; q += stride) {
    const unsigned long long gp = (unsigned long long)(p.path0 + q);
    int s = __ldg(p.starts + (long long)(gp % (unsigned long long)p.nstarts));
    float tf = (float)p.V[s], td = tf;
    float* fo = p.fit + (size_t)q * p.nsteps;
    float* go = p.dir + (size_t)q * p.nsteps;
    for (int t = 0; t < p.nsteps; t++) {
      const float4 r0 = __ldg(p.table + 2 * (size_t)s), r1 = __ldg(p.table + 2 * (size_t)s + 1);
      const float kl = __ldg(p.kl + s);
      float u0;
      bool slip;
      int aslip;
      if (INJ) {
        const float4 u = *reinterpret_cast<const float4*>(p.injected + ((size_t)q * (size_t)p.injected_steps + t) * 4);
        u0 = u.x; slip = (double)u.z < p.rho; aslip = min((int)(4.0f * u.w), 3);
      } else {
        const uint4 r = philox4x32_10(p.keys, (uint32_t)t, 0u, (uint32_t)gp, (uint32_t)(gp >> 32));
        u0 = u01(r.x); slip = (r.z >> 8) < p.rho_k; aslip = (int)(r.w >> 30);
      }
      const bool b0 = r0.x < u0, b1 = r0.y < u0, b2 = r0.z < u0;  // choose_action: the cw < t walk
      const int a = (int)b0 + (int)b1 + (int)b2;
      float bf = r1.x;
      bf = b0 ? r1.y : bf; bf = b1 ? r1.z : bf; bf = b2 ? r1.w : bf;
      const int aa = slip ? aslip : a;                             // propagator   GridWorld.jl:113-115
      const unsigned blocked = (__float_as_uint(r0.w) >> aa) & 1u;
      int d = (aa & 2) ? p.H : 1;
      d = (aa & 1) ? -d : d;
      const int s1 = blocked ? s : s + d;
      tf = __fadd_rn(tf, bf);
      const float rw = s1 == p.goal ? p.rwd_goal : p.rwd_step;     // reward_function(s, a, s')   GridWorld.jl:116
      td = __fadd_rn(td, __fsub_rn(rw, kl));
      fo[t] = tf;
      go[t] = td;
      if (p.trace) { int* tr = p.trace + ((size_t)q * p.nsteps + t) * 3; tr[0] = s; tr[1] = a; tr[2] = s1; }
      s = s1;
    }
  }
}

void launch_path_estimates(const PathEstArgs& a, int num_sms, cudaStream_t st) {
  if (a.npaths <= 0) return;
  const long long want = (a.npaths + 255) / 256, cap = (long long)num_sms * 8;
  const int grid = (int)(want < cap ? want : cap);
  if (a.injected) path_estimates_kernel<true><<<grid, 256, 0, st>>>(a);
  else path_estimates_kernel<false><<<grid, 256, 0, st>>>(a);
  g_launches++;
}

// ---- rollout ---------------------------------------------------------------------------------------
__device__ __forceinline__ float logaddexpf_dev(float x, float y) {
  if (x == -CUDART_INF_F) return y;
  if (y == -CUDART_INF_F) return x;
  return fmaxf(x, y) + log1pf(expf(-fabsf(x - y)));
}

constexpr int ROLL_THREADS = 256;

// STAGED: the whole policy table (<= ROLL_STAGE_MAX_BYTES: the reference-scale problems) is copied into shared memory
// by one bulk asynchronous copy per CTA and every step reads its record with two LDS.128 instead of a global load.
constexpr size_t ROLL_STAGE_MAX_BYTES = 24 * 1024;  // x 8 resident CTAs per SM


// Tried and dropped (measured on B200, profiles/r2_roll_tiled_check.txt): a second copy of the records in 8 x 8 blocks of
// cells so that a horizontal step stays within a DRAM page -- bit-identical records, 12 % slower at 4098^2 (1.51e11 against
// 1.71e11 transitions/s): the extra address arithmetic costs issue slots and the sector traffic does not change.
// REC16: the 16-byte records of policy_table_kernel (grids whose 32-byte table does not fit L2: half the footprint, two
// cells per DRAM sector, eight per line); decoded with five extra integer operations per step.
template <int EST, bool INJ, bool TRACE, bool STAGED, bool REC16 = false>
__global__ void __launch_bounds__(ROLL_THREADS)
rollout_kernel(const RolloutArgs p) {
  extern __shared__ __align__(128) unsigned char roll_smem[];
  __shared__ __align__(8) unsigned long long stage_bar;
  if (STAGED) {
    if (threadIdx.x == 0) {
      mbar_init(&stage_bar, 1);
      mbar_expect_tx(&stage_bar, (unsigned)p.table_bytes);
      bulk_load_1d(roll_smem, p.table, (unsigned)p.table_bytes, &stage_bar);
    }
    __syncthreads();
    mbar_wait(&stage_bar, 0);
  }
  const float4* const stab = reinterpret_cast<const float4*>(roll_smem);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long nwarps = (long long)gridDim.x * (ROLL_THREADS / 32);
  const long long wid = (long long)blockIdx.x * (ROLL_THREADS / 32) + (threadIdx.x >> 5);
  // contiguous range of local path indices for this warp: [wbase, wbase + wcount)
  const long long per = (p.npaths + nwarps - 1) / nwarps;
  const long long wbase = wid * per;
  if (wbase >= p.npaths) return;  // warp-uniform
  const uint32_t wcount = (uint32_t)min(per, p.npaths - wbase);
  const unsigned long long gbase = (unsigned long long)(p.path0 + wbase);  // global id of the warp's first path
  const uint32_t nstarts = (uint32_t)p.nstarts;
  // start slot of the next path to hand out, kept incrementally (one 64-bit division per warp, not per path)
  uint32_t slot_next = (uint32_t)(gbase % (unsigned long long)p.nstarts);
  uint32_t next = 0;  // paths of this warp handed out so far
  const uint32_t hor = (uint32_t)min(INJ ? min(p.horizon, p.injected_steps) : p.horizon, 0xffffffffLL);
  const int H = p.H, goal = p.goal;
  const double gam = p.gamma, rho = p.rho;  // injected uniforms only (the Philox path tests the 24-bit integers)

  bool active = false;
  uint32_t q = 0;             // path index within the warp's range
  uint32_t plo = 0, phi = 0;  // global path id (Philox counter words 2, 3)
  int s = 0;
  uint32_t t = 0, a0 = 0;
  float logw = 0.f, acc = -CUDART_INF_F;

  for (;;) {
    const unsigned need = __ballot_sync(FULL, !active);
    if (need) {
      const uint32_t rank = __popc(need & lt_mask);
      if (!active && next + rank < wcount) {
        q = next + rank;
        const unsigned long long gp = gbase + q;
        plo = (uint32_t)gp; phi = (uint32_t)(gp >> 32);
        uint32_t slot = slot_next + rank;
        if (nstarts >= 32u) { if (slot >= nstarts) slot -= nstarts; }
        else slot %= nstarts;
        s = __ldg(p.starts + slot);
        t = 0; a0 = 0; logw = 0.f; acc = -CUDART_INF_F;
        active = true;
      }
      // saturating: once the warp's range is handed out, `next` stays at wcount however long the last lanes walk on
      // (estimator B with a long horizon would otherwise wrap the 32-bit counter and hand finished paths out again)
      if (next < wcount) {
        const uint32_t handed = min((uint32_t)__popc(need), wcount - next);
        next += handed;
        slot_next += handed;
        if (nstarts >= 32u) { if (slot_next >= nstarts) slot_next -= nstarts; }
        else slot_next %= nstarts;
      }
      if (!__any_sync(FULL, active)) break;
    }
    if (active) {
      bool done;
      if (s == goal) {
        // Z(goal) = exp(0 - V(goal))   value_iteration.jl:93-94
        if (EST == 0) logw = __fsub_rn(logw, p.vgoal);
        else acc = logaddexpf_dev(acc, __fadd_rn(__fmul_rn((float)t, p.lgam), __fsub_rn(logw, p.vgoal)));
        done = true;
      } else if (t >= hor) {
        done = true;
      } else {
        // the whole 32-byte record in one 256-bit load (LDG.E.256, sm_100): one L1 wavefront per distinct line instead of
        // two passes over the same sector; issued first (volatile keeps it there) so that the ~45 instructions of Philox
        // run under the load's latency -- at 4096^2 half of these loads miss L2
        float4 r0, r1;
        if (REC16) {
          uint4 q16;
          asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q16.x), "=r"(q16.y), "=r"(q16.z), "=r"(q16.w)
                       : "l"(p.table16 + s));
          r0 = make_float4(__uint_as_float(q16.x & 0x3fffffffu), __uint_as_float(q16.y & 0x3fffffffu), __uint_as_float(q16.z),
                           __uint_as_float((q16.x >> 30) | ((q16.y >> 30) << 2)));
          const float bf0 = __uint_as_float(q16.w);
          r1 = make_float4(bf0, bf0, bf0, bf0);
        } else if (STAGED) {
          r0 = stab[2 * s];
          r1 = stab[2 * s + 1];
        } else {
          const float4* rec = p.table + 2 * (size_t)s;
          asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=f"(r0.x), "=f"(r0.y), "=f"(r0.z), "=f"(r0.w), "=f"(r1.x), "=f"(r1.y), "=f"(r1.z), "=f"(r1.w)
                       : "l"(rec));
        }
        float u0;
        bool stop, slip;
        int aslip;
        if (INJ) {
          const float4 u = *reinterpret_cast<const float4*>(
              p.injected + (((size_t)wbase + q) * (size_t)p.injected_steps + t) * 4);
          u0 = u.x;
          stop = (double)u.y >= gam;
          slip = (double)u.z < rho;
          aslip = min((int)(4.0f * u.w), 3);
        } else {
          // u = k * 2^-24 with the 24-bit integer k = word >> 8, so the tests u >= gamma and u < randomness (against the
          // problem's Float64 gamma / randomness) are exactly k >= ceil(gamma * 2^24) and k < ceil(randomness * 2^24)
          const uint4 r = philox4x32_10(p.keys, t, 0u, plo, phi);
          u0 = u01(r.x);
          stop = (r.y >> 8) >= p.gamma_k;
          slip = (r.z >> 8) < p.rho_k;
          aslip = (int)(r.w >> 30);
        }
        // StatsBase sample(a, Weights): i = 1; cw = w[1]; while cw < t && i < n: i += 1; cw += w[i]
        // c0 <= c1 <= c2, so the walk stops at the number of running sums below u0 (branch-free).
        const bool b0 = r0.x < u0, b1 = r0.y < u0, b2 = r0.z < u0;
        const int a = (int)b0 + (int)b1 + (int)b2;
        float bf = r1.x;
        bf = b0 ? r1.y : bf;
        bf = b1 ? r1.z : bf;
        bf = b2 ? r1.w : bf;
        logw = __fadd_rn(logw, bf);
        a0 = (t == 0u) ? (uint32_t)a : a0;
        if (EST == 1) acc = logaddexpf_dev(acc, __fadd_rn(__fadd_rn(__fmul_rn((float)t, p.lgam), p.l1mg), logw));
        // estimator A: continue w.p. gamma, else take the (1 - gamma) * 1 branch of z_recursion and stop
        done = (EST == 0) && stop;
        if (!done) {
          // propagator: take_step(s, ifelse(rand() < randomness, sample(0:3), a))   GridWorld.jl:113-115
          const int aa = slip ? aslip : a;
          const unsigned blocked = (__float_as_uint(r0.w) >> aa) & 1u;
          int d = (aa & 2) ? H : 1;
          d = (aa & 1) ? -d : d;
          d = blocked ? 0 : d;
          if (TRACE) {
            if ((long long)t < p.trace_steps) {
              int* tr = p.trace + (((size_t)wbase + q) * (size_t)p.trace_steps + t) * 3;
              tr[0] = s; tr[1] = a; tr[2] = s + d;
            }
          }
          s += d;
          t++;  // t counts sampled transitions: it is also the path's length
        }
      }
      if (done) {
        // the path's statistics leave in one 8-byte store (STG.64)
        p.rec[(size_t)wbase + q] = make_uint2(__float_as_uint((EST == 0) ? logw : acc), a0 | (t << 2));
        active = false;
      }
    }
  }
}

void launch_rollout(const RolloutArgs& a, int num_sms, cudaStream_t st) {
  if (a.npaths <= 0) return;
  // persistent grid: 8 CTAs of 256 threads per SM (2048 resident threads), fewer for small jobs
  long long want = (a.npaths + ROLL_THREADS - 1) / ROLL_THREADS;
  long long cap = (long long)num_sms * 8;
  const int grid = (int)(want < cap ? want : cap);
  const bool inj = a.injected != nullptr, tr = a.trace != nullptr;
#define PIQL_ROLL(E, I, T) rollout_kernel<E, I, T, false><<<grid, ROLL_THREADS, 0, st>>>(a)
  static const bool no_stage = getenv("PIQL_ROLL_NO_STAGE") != nullptr;  // A/B timing only
  if (!no_stage && !inj && !tr && a.table_bytes > 0 && a.table_bytes <= (long long)ROLL_STAGE_MAX_BYTES) {
    // reference-scale tables: staged in shared memory
    if (a.estimator == 0) rollout_kernel<0, false, false, true><<<grid, ROLL_THREADS, (size_t)a.table_bytes, st>>>(a);
    else rollout_kernel<1, false, false, true><<<grid, ROLL_THREADS, (size_t)a.table_bytes, st>>>(a);
  } else if (a.table16 && !inj && !tr) {
    if (a.estimator == 0) rollout_kernel<0, false, false, false, true><<<grid, ROLL_THREADS, 0, st>>>(a);
    else rollout_kernel<1, false, false, false, true><<<grid, ROLL_THREADS, 0, st>>>(a);
  } else if (a.estimator == 0) {
    if (inj) { if (tr) PIQL_ROLL(0, true, true); else PIQL_ROLL(0, true, false); }
    else { if (tr) PIQL_ROLL(0, false, true); else PIQL_ROLL(0, false, false); }
  } else {
    if (inj) { if (tr) PIQL_ROLL(1, true, true); else PIQL_ROLL(1, true, false); }
    else { if (tr) PIQL_ROLL(1, false, true); else PIQL_ROLL(1, false, false); }
  }
#undef PIQL_ROLL
  g_launches++;
}

// ---- log-sum-exp reduction of path records ----------------------------------------------------------
// Online (max, sum) update with one exponential per record.
// The exponentials are fp32 (one MUFU each); the running sum is kept in double so that a bucket
// of 10^7 records loses nothing to accumulation order.
struct Lse {
  float m;
  double s;
  uint32_t c;
  __device__ __forceinline__ void init() { m = -CUDART_INF_F; s = 0.0; c = 0u; }
  __device__ __forceinline__ void add(float x) {
    if (x > m) { s = s * (double)expf(m - x) + 1.0; m = x; }
    else if (x != -CUDART_INF_F) s += (double)expf(x - m);
    c++;
  }
  __device__ __forceinline__ void merge(float m2, double s2, uint32_t c2) {
    if (m2 > m) { s = s * (double)expf(m - m2) + s2; m = m2; }
    else if (m2 != -CUDART_INF_F) s += s2 * (double)expf(m2 - m);
    c += c2;
  }
};

// Regime "many start slots": one thread per slot walks that slot's records (stride nstarts), so
// consecutive threads read consecutive records; no cross-thread combine is needed at all.
__global__ void __launch_bounds__(256)
reduce_by_slot_kernel(const uint2* __restrict__ rec, long long path0,
                      long long npaths, long long nstarts, float* __restrict__ acc_max, float* __restrict__ acc_sum,
                      uint32_t* __restrict__ acc_cnt, unsigned long long* __restrict__ transitions) {
  __shared__ unsigned long long red[32];
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long tr = 0;
  if (slot < nstarts) {
    Lse b[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) b[a].init();
    // first local index q with (path0 + q) % nstarts == slot
    long long q = (slot - path0 % nstarts + nstarts) % nstarts;
    for (; q < npaths; q += nstarts) {
      const uint2 rc = __ldg(rec + q);
      const float x = __uint_as_float(rc.x);
      const uint32_t me = rc.y;
      tr += me >> 2;
      const uint32_t a = me & 3u;
#pragma unroll
      for (int k = 0; k < NA; k++) if (a == (uint32_t)k) b[k].add(x);
    }
#pragma unroll
    for (int a = 0; a < NA; a++) {
      if (b[a].c == 0) continue;
      const size_t i = (size_t)slot * NA + a;
      Lse g; g.m = acc_max[i]; g.s = acc_sum[i]; g.c = acc_cnt[i];
      g.merge(b[a].m, b[a].s, b[a].c);
      acc_max[i] = g.m; acc_sum[i] = (float)g.s; acc_cnt[i] = g.c;
    }
  }
  const unsigned long long r = block_sum(tr, red);
  if (threadIdx.x == 0 && r) atomicAdd(transitions, r);
}

// Regime "few start slots": grid.y = slot, grid.x = chunks of that slot's records; warp-shuffle
// then block-level combine of (max, sum, count) per action, one partial per block, merged in a
// fixed order by reduce_merge_kernel.
__global__ void __launch_bounds__(256)
reduce_block_kernel(const uint2* __restrict__ rec, long long path0,
                    long long npaths, long long nstarts, float* __restrict__ part /*[slot][chunk][a][3]*/,
                    unsigned long long* __restrict__ transitions) {
  __shared__ float sm_m[NA][8];
  __shared__ double sm_s[NA][8];
  __shared__ uint32_t sm_c[NA][8];
  __shared__ unsigned long long red[32];
  const long long slot = blockIdx.y;
  const long long q0 = (slot - path0 % nstarts + nstarts) % nstarts;
  const long long nrec = q0 < npaths ? (npaths - q0 + nstarts - 1) / nstarts : 0;  // records of this slot
  const long long per = (nrec + gridDim.x - 1) / gridDim.x;
  const long long jb = (long long)blockIdx.x * per, je = min(nrec, jb + per);
  Lse b[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) b[a].init();
  unsigned long long tr = 0;
  for (long long j = jb + threadIdx.x; j < je; j += blockDim.x) {
    const long long q = q0 + j * nstarts;
    const uint2 rc = __ldg(rec + q);
    const float x = __uint_as_float(rc.x);
    const uint32_t me = rc.y;
    tr += me >> 2;
    const uint32_t a = me & 3u;
#pragma unroll
    for (int k = 0; k < NA; k++) if (a == (uint32_t)k) b[k].add(x);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NA; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(FULL, b[a].m, o);
      const double s2 = __shfl_xor_sync(FULL, b[a].s, o);
      const uint32_t c2 = __shfl_xor_sync(FULL, b[a].c, o);
      // symmetric combine: both partners end up with bit-identical (max, sum, count)
      const float m = fmaxf(b[a].m, m2);
      const double sa = (b[a].m == -CUDART_INF_F) ? 0.0 : b[a].s * (double)expf(b[a].m - m);
      const double sb = (m2 == -CUDART_INF_F) ? 0.0 : s2 * (double)expf(m2 - m);
      b[a].s = sa + sb;
      b[a].m = m; b[a].c += c2;
    }
    if (lane == 0) { sm_m[a][warp] = b[a].m; sm_s[a][warp] = b[a].s; sm_c[a][warp] = b[a].c; }
  }
  __syncthreads();
  if (threadIdx.x < NA) {
    const int a = threadIdx.x;
    Lse g; g.init();
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) g.merge(sm_m[a][w], sm_s[a][w], sm_c[a][w]);
    float* o = part + (((size_t)slot * gridDim.x + blockIdx.x) * NA + a) * 3;
    o[0] = g.m; o[1] = (float)g.s; o[2] = __uint_as_float(g.c);
  }
  const unsigned long long r = block_sum(tr, red);
  if (threadIdx.x == 0 && r) atomicAdd(transitions, r);
}
__global__ void reduce_merge_kernel(const float* __restrict__ part, int chunks, long long nstarts,
                                    float* __restrict__ acc_max, float* __restrict__ acc_sum, uint32_t* __restrict__ acc_cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // slot*4 + a
  if (i >= nstarts * NA) return;
  const long long slot = i / NA;
  const int a = (int)(i % NA);
  Lse g; g.m = acc_max[i]; g.s = acc_sum[i]; g.c = acc_cnt[i];
  for (int c = 0; c < chunks; c++) {
    const float* o = part + (((size_t)slot * chunks + c) * NA + a) * 3;
    g.merge(o[0], o[1], __float_as_uint(o[2]));
  }
  acc_max[i] = g.m; acc_sum[i] = (float)g.s; acc_cnt[i] = g.c;
}

constexpr long long REDUCE_FEW_SLOTS = 2048;
static int reduce_chunks(long long npaths, long long nstarts, int num_sms) {
  const long long per_slot = (npaths + nstarts - 1) / nstarts;
  long long c = (per_slot + 256 * 8 - 1) / (256 * 8);          // >= 8 records per thread
  const long long cap = (4LL * num_sms + nstarts - 1) / nstarts; // ~4 CTAs per SM in total
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return (int)c;
}
size_t reduce_scratch_floats(long long npaths, long long nstarts) {
  if (nstarts > REDUCE_FEW_SLOTS) return 0;
  return (size_t)nstarts * reduce_chunks(npaths, nstarts, 148) * NA * 3;
}
void launch_reduce(const uint2* rec, long long path0, long long npaths, long long nstarts,
                   float* acc_max, float* acc_sum, uint32_t* acc_cnt, float* scratch, size_t scratch_floats,
                   unsigned long long* transitions, int num_sms, cudaStream_t st) {
  if (npaths <= 0) return;
  if (nstarts > REDUCE_FEW_SLOTS) {
    reduce_by_slot_kernel<<<(unsigned)((nstarts + 255) / 256), 256, 0, st>>>(rec, path0, npaths, nstarts, acc_max,
                                                                             acc_sum, acc_cnt, transitions);
    g_launches++;
  } else {
    (void)num_sms;
    const int chunks = reduce_chunks(npaths, nstarts, 148);
    (void)scratch_floats;
    dim3 grid(chunks, (unsigned)nstarts);
    reduce_block_kernel<<<grid, 256, 0, st>>>(rec, path0, npaths, nstarts, scratch, transitions);
    reduce_merge_kernel<<<(unsigned)((nstarts * NA + 127) / 128), 128, 0, st>>>(scratch, chunks, nstarts, acc_max, acc_sum, acc_cnt);
    g_launches += 2;
  }
}

__global__ void split_records_kernel(const uint2* __restrict__ rec, long long n, float* __restrict__ logw,
                                     uint32_t* __restrict__ meta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 r = rec[i];
  if (logw) logw[i] = __uint_as_float(r.x);
  if (meta) meta[i] = r.y;
}
void launch_split_records(const uint2* rec, long long n, float* logw, uint32_t* meta, cudaStream_t st) {
  if (n <= 0) return;
  split_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec, n, logw, meta);
  g_launches++;
}

// ---- accumulator maintenance -----------------------------------------------------------------------
__global__ void mc_reset_kernel(float* m, float* s, uint32_t* c, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { m[i] = -CUDART_INF_F; s[i] = 0.f; c[i] = 0u; }
}
void launch_mc_reset(float* acc_max, float* acc_sum, uint32_t* acc_cnt, long long n, cudaStream_t st) {
  mc_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc_max, acc_sum, acc_cnt, n);
  g_launches++;
}
// counters[0]: transitions since the last reset, counters[1]: running total over resets
__global__ void fold_counter_kernel(unsigned long long* counters) {
  counters[1] += counters[0];
  counters[0] = 0ull;
}
void launch_fold_counter(unsigned long long* counters, cudaStream_t st) {
  fold_counter_kernel<<<1, 1, 0, st>>>(counters);
  g_launches++;
}
// After an all-reduce(MAX) of lse_max across shards: bring the local sum to the global shift so
// that an all-reduce(SUM) of lse_sum finishes the log-sum-exp.
__global__ void mc_rescale_kernel(float* m, float* s, const float* __restrict__ gmax, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float g = gmax[i], l = m[i];
    s[i] = (l == -CUDART_INF_F) ? 0.f : s[i] * expf(l - g);
    m[i] = g;
  }
}
void launch_mc_rescale(float* acc_max, float* acc_sum, const float* gmax, long long n, cudaStream_t st) {
  mc_rescale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc_max, acc_sum, gmax, n);
  g_launches++;
}
// TF / TC: output types -- double / uint64 (the ABI's default) or float / uint32 (half the device-to-host bytes)
template <typename TF, typename TC>
__global__ void mc_finalize_kernel(const float* __restrict__ m, const float* __restrict__ s, const uint32_t* __restrict__ c,
                                   const int* __restrict__ starts, long long nstarts, TF* __restrict__ logz_sa,
                                   TC* __restrict__ count_sa, TF* __restrict__ logz_s) {
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nstarts) return;
  const size_t cellidx = (size_t)starts[slot];
  double mm = -CUDART_INF;
  unsigned long long ctot = 0;
#pragma unroll
  for (int a = 0; a < NA; a++) {
    const size_t i = (size_t)slot * NA + a;
    const uint32_t cnt = c[i];
    if (logz_sa) logz_sa[cellidx * NA + a] = (TF)(cnt ? (double)m[i] + log((double)s[i]) - log((double)cnt) : CUDART_NAN);
    if (count_sa) count_sa[cellidx * NA + a] = (TC)cnt;
    if (cnt && (double)m[i] > mm) mm = (double)m[i];
    ctot += cnt;
  }
  if (logz_s) {
    double tot = 0.0;
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const size_t i = (size_t)slot * NA + a;
      if (c[i]) tot += (double)s[i] * exp((double)m[i] - mm);
    }
    logz_s[cellidx] = (TF)(ctot ? mm + log(tot) - log((double)ctot) : CUDART_NAN);
  }
}
void launch_mc_finalize(const float* acc_max, const float* acc_sum, const uint32_t* acc_cnt, const int* starts,
                        long long nstarts, double* logz_sa, unsigned long long* count_sa, double* logz_s,
                        cudaStream_t st) {
  mc_finalize_kernel<double, unsigned long long><<<(unsigned)((nstarts + 255) / 256), 256, 0, st>>>(
      acc_max, acc_sum, acc_cnt, starts, nstarts, logz_sa, count_sa, logz_s);
  g_launches++;
}
void launch_mc_finalize_f32(const float* acc_max, const float* acc_sum, const uint32_t* acc_cnt, const int* starts,
                            long long nstarts, float* logz_sa, uint32_t* count_sa, float* logz_s, cudaStream_t st) {
  mc_finalize_kernel<float, uint32_t><<<(unsigned)((nstarts + 255) / 256), 256, 0, st>>>(acc_max, acc_sum, acc_cnt, starts,
                                                                                         nstarts, logz_sa, count_sa, logz_s);
  g_launches++;
}
__global__ void fill_f32_kernel(float* dst, float v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
void launch_fill_f32(float* dst, float v, long long n, cudaStream_t st) {
  fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
  g_launches++;
}
// The Monte-Carlo form of logz_updateV (value_iteration.jl:119-124): V(s) += log Z^(s) / beta for every start slot that
// received paths, log Z^(s) pooling the slot's four first-action buckets.
__global__ void mc_apply_update_kernel(const float* __restrict__ m, const float* __restrict__ s, const uint32_t* __restrict__ c,
                                       const int* __restrict__ starts, long long nstarts, double beta, double* __restrict__ V) {
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nstarts) return;
  double mm = -CUDART_INF;
  unsigned long long ctot = 0;
#pragma unroll
  for (int a = 0; a < NA; a++) {
    const size_t i = (size_t)slot * NA + a;
    if (c[i] && (double)m[i] > mm) mm = (double)m[i];
    ctot += c[i];
  }
  if (!ctot) return;
  double tot = 0.0;
#pragma unroll
  for (int a = 0; a < NA; a++) {
    const size_t i = (size_t)slot * NA + a;
    if (c[i]) tot += (double)s[i] * exp((double)m[i] - mm);
  }
  V[starts[slot]] += (mm + log(tot) - log((double)ctot)) / beta;
}
void launch_mc_apply_update(const float* acc_max, const float* acc_sum, const uint32_t* acc_cnt, const int* starts,
                            long long nstarts, double beta, double* V, cudaStream_t st) {
  mc_apply_update_kernel<<<(unsigned)((nstarts + 255) / 256), 256, 0, st>>>(acc_max, acc_sum, acc_cnt, starts, nstarts, beta, V);
  g_launches++;
}
__global__ void fill_f64_kernel(double* dst, double v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
void launch_fill_f64(double* dst, double v, long long n, cudaStream_t st) {
  fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
  g_launches++;
}

}  // namespace piql
