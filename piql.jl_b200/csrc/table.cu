// table.cu -- soft value iteration on a tabulated ControlProblem (any finite problem behind PIQL's closure interface,
// src/PIQL.jl:12-23): states 0..ns-1, actions 0..na-1, CSR over (state, action) of (successor, probability, reward).
//   propagator_average(s,a,f) = sum_e p_e f(s'_e)        logpropagatorexp(s,a,f) = log sum_e p_e exp f(s'_e)
// The same recursions as backup.cu (truevalue_recursion src/PIQL.jl:95-101, logz_recursion :125-130,
// optimality_gap_recursion src/value_iteration.jl:160-169) with the V-determined Q (s,a) -> controlV(ctrl,V,s,a)
// (value_iteration.jl:97,157), Jacobi sweeps in Float64, the reference's stop rule sum |new - old| < tol.
// This is the path for problems that are not the implicit gridworld (BinaryBandit, src/Problems/BinaryBandit.jl, or any
// user ControlProblem the host tabulates); one thread per state walks its CSR rows, values gathered through L2.
#include "internal.h"

namespace piql {

constexpr int TAB_MAX_ACTIONS = 16;
constexpr int TAB_THREADS = 128;

__device__ __forceinline__ double tab_pa(const TableDev& t, long long row, const double* __restrict__ f) {
  double out = 0.0;  // propagator_average
  for (long long e = t.ptr[row]; e < t.ptr[row + 1]; e++) out += t.prob[e] * f[t.succ[e]];
  return out;
}
__device__ __forceinline__ double tab_rbar(const TableDev& t, long long row, double scale) {
  double out = 0.0;  // average_reward  PIQL.jl:55-57
  for (long long e = t.ptr[row]; e < t.ptr[row + 1]; e++) out += t.prob[e] * (scale * t.rew[e]);
  return out;
}
__device__ __forceinline__ double tab_log_pa(const TableDev& t, long long row, const double* __restrict__ f) {
  double m = -CUDART_INF;  // logpropagatorexp
  for (long long e = t.ptr[row]; e < t.ptr[row + 1]; e++) m = fmax(m, log(t.prob[e]) + f[t.succ[e]]);
  if (m == -CUDART_INF) return m;
  double acc = 0.0;
  for (long long e = t.ptr[row]; e < t.ptr[row + 1]; e++) acc += exp(log(t.prob[e]) + f[t.succ[e]] - m);
  return m + log(acc);
}
__device__ __forceinline__ double tab_lse(const double* x, int n) {
  double m = -CUDART_INF;
  for (int i = 0; i < n; i++) m = fmax(m, x[i]);
  if (m == -CUDART_INF) return m;
  double s = 0.0;
  for (int i = 0; i < n; i++) s += exp(x[i] - m);
  return m + log(s);
}

// Per-(state, action) constants of a logz / gap solve: logz A = log prior + beta cV + (1-beta) cQ; gap pi (and D per state).
__global__ void __launch_bounds__(TAB_THREADS)
table_prep_kernel(TableDev t, int kind, double scale, double beta, const double* __restrict__ V, const double* __restrict__ nu,
                  const double* __restrict__ Q, double* __restrict__ aux, double* __restrict__ aux1) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= t.ns) return;
  const int na = t.na;
  double cv[TAB_MAX_ACTIONS], q[TAB_MAX_ACTIONS], x[TAB_MAX_ACTIONS];
  for (int a = 0; a < na; a++) {
    cv[a] = t.gamma * tab_pa(t, s * na + a, V) - V[s] + tab_rbar(t, s * na + a, scale);  // controlV  PIQL.jl:63-66
    q[a] = Q ? Q[s * na + a] : cv[a];       // a tabulated Q, or (s,a) -> controlV(ctrl,V,s,a)   value_iteration.jl:97,157
    x[a] = t.lprior[s * na + a] + q[a];
  }
  const double norm = tab_lse(x, na);  // Qnormalization  PIQL.jl:72-75
  double D = 0.0;
  for (int a = 0; a < na; a++) {
    const double cq = q[a] - norm;  // controlQ  PIQL.jl:79-81
    if (kind == KIND_LOGZ) {
      aux[s * na + a] = t.lprior[s * na + a] + beta * cv[a] + (1.0 - beta) * cq;
    } else {
      const double pi = exp(t.lprior[s * na + a] + cq);  // action_probability  PIQL.jl:83-85
      const double cvn = t.gamma * tab_pa(t, s * na + a, nu) - nu[s] + tab_rbar(t, s * na + a, scale);
      aux[s * na + a] = pi;
      D += pi * (cq - cvn);
    }
  }
  if (kind == KIND_GAP) aux1[s] = D;
}

template <int KIND>
__global__ void __launch_bounds__(TAB_THREADS)
table_sweep_kernel(TableDev t, double scale, const double* __restrict__ X, double* __restrict__ Y, const double* __restrict__ V,
                   const double* __restrict__ aux, const double* __restrict__ aux1, double* __restrict__ partial) {
  __shared__ double red[32];
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  if (s < t.ns) {
    const int na = t.na;
    const double c = X[s];
    double nw;
    if (t.term[s]) {
      nw = KIND == KIND_LOGZ ? 0.0 - V[s] : 0.0;  // value_iteration.jl:66-67,93-94,133-134
    } else if (KIND == KIND_NU) {
      double x[TAB_MAX_ACTIONS];
      for (int a = 0; a < na; a++)
        x[a] = t.lprior[s * na + a] + t.gamma * tab_pa(t, s * na + a, X) + tab_rbar(t, s * na + a, scale);
      nw = tab_lse(x, na);
    } else if (KIND == KIND_LOGZ) {
      double x[TAB_MAX_ACTIONS];
      for (int a = 0; a < na; a++)
        x[a] = aux[s * na + a] + logaddexp(t.lgamma + tab_log_pa(t, s * na + a, X), t.l1mgamma);
      nw = tab_lse(x, na);
    } else {
      nw = aux1[s];
      for (int a = 0; a < na; a++) nw += aux[s * na + a] * tab_pa(t, s * na + a, X);  // no gamma  value_iteration.jl:165
    }
    d = fabs(nw - c);
    Y[s] = KIND == KIND_LOGZ ? logaddexp(-0.6931471805599453 + nw, -0.6931471805599453 + c) : nw;  // eta = 1/2  :101-102
  }
  if (partial) {
    const double r = block_sum(d, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
  }
}

__global__ void table_update_v_kernel(long long ns, const double* __restrict__ V, const double* __restrict__ logz, double beta,
                                      double* __restrict__ out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < ns) out[s] = V[s] + logz[s] / beta;  // logz_updateV  value_iteration.jl:119-124
}

static inline unsigned tblk(long long n) { return (unsigned)((n + TAB_THREADS - 1) / TAB_THREADS); }

void launch_table_prep(const TableDev& t, int kind, double scale, double beta, const double* V, const double* nu, double* aux,
                       double* aux1, cudaStream_t st, const double* Q) {
  table_prep_kernel<<<tblk(t.ns), TAB_THREADS, 0, st>>>(t, kind, scale, beta, V, nu, Q, aux, aux1);
  g_launches++;
}
void launch_table_sweep(const TableDev& t, int kind, double scale, const double* X, double* Y, const double* V, const double* aux,
                        const double* aux1, double* partial, double* residual, cudaStream_t st) {
  const unsigned nb = tblk(t.ns);
  switch (kind) {
    case KIND_NU: table_sweep_kernel<KIND_NU><<<nb, TAB_THREADS, 0, st>>>(t, scale, X, Y, V, aux, aux1, partial); break;
    case KIND_LOGZ: table_sweep_kernel<KIND_LOGZ><<<nb, TAB_THREADS, 0, st>>>(t, scale, X, Y, V, aux, aux1, partial); break;
    default: table_sweep_kernel<KIND_GAP><<<nb, TAB_THREADS, 0, st>>>(t, scale, X, Y, V, aux, aux1, partial); break;
  }
  g_launches++;
  if (partial) launch_residual_final(partial, (int)nb, residual, st);
}
void launch_table_update_v(long long ns, const double* V, const double* logz, double beta, double* out, cudaStream_t st) {
  table_update_v_kernel<<<tblk(ns), TAB_THREADS, 0, st>>>(ns, V, logz, beta, out);
  g_launches++;
}
int table_blocks(long long ns) { return (int)tblk(ns); }

}  // namespace piql

// ================================================================================================================
// Monte-Carlo path sampling on a tabulated problem: the estimator of DESIGN.md section 3 (sampled stop) with the table's
// own transition law.  Per step of path p the sampler consumes three uniforms of one Philox4x32-10 call, counter
// (step, 0, p_lo, p_hi): u0 -> action (StatsBase walk over the running action probabilities), u1 -> stop test u1 >= gamma,
// u2 -> successor (the same walk over the running transition probabilities of row (s, a)); at a terminal state
// logw -= V(s) and the path ends (Z(terminal) = exp(-V), value_iteration.jl:93-94).
// ================================================================================================================
namespace piql {

// per-(state, action) sampler tables from V: running normalised action probabilities and beta * fitness (PIQL.jl:83-89)
__global__ void __launch_bounds__(TAB_THREADS)
table_policy_kernel(TableDev t, double scale, double beta, const double* __restrict__ V, const double* __restrict__ Q,
                    float* __restrict__ cdf, float* __restrict__ bf, float* __restrict__ vterm) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= t.ns) return;
  const int na = t.na;
  vterm[s] = (float)V[s];
  double cv[TAB_MAX_ACTIONS], q[TAB_MAX_ACTIONS], x[TAB_MAX_ACTIONS], w[TAB_MAX_ACTIONS];
  for (int a = 0; a < na; a++) {
    cv[a] = t.term[s] ? 0.0 : t.gamma * tab_pa(t, s * na + a, V) - V[s] + tab_rbar(t, s * na + a, scale);
    q[a] = (Q && !t.term[s]) ? Q[s * na + a] : cv[a];
    x[a] = t.lprior[s * na + a] + q[a];
  }
  const double norm = tab_lse(x, na);
  double sum = 0.0, cw = 0.0;
  for (int a = 0; a < na; a++) { w[a] = exp(t.lprior[s * na + a] + q[a] - norm); sum += w[a]; }
  for (int a = 0; a < na; a++) {
    cw += w[a];
    cdf[s * na + a] = (float)(cw / sum);
    bf[s * na + a] = (float)(beta * (cv[a] - (q[a] - norm)));      // beta * fitness   PIQL.jl:87-89
  }
}
// running transition probabilities within each (state, action) row
__global__ void __launch_bounds__(TAB_THREADS) table_cum_kernel(TableDev t, float* __restrict__ cum) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= t.ns * t.na) return;
  double cw = 0.0;
  for (long long e = t.ptr[row]; e < t.ptr[row + 1]; e++) { cw += t.prob[e]; cum[e] = (float)cw; }
}

struct TableRollArgs {
  TableDev t;
  const float* cdf; const float* bf; const float* cum; const float* vterm;
  uint32_t gamma_k;
  long long horizon, path0, npaths, nstarts;
  const int* starts;
  PhiloxKeys keys;
  float* logw; uint32_t* meta;
};

__global__ void __launch_bounds__(256) table_rollout_kernel(const TableRollArgs p) {
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  const long long nwarps = (long long)gridDim.x * 8, wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long per = (p.npaths + nwarps - 1) / nwarps, wbase = wid * per;
  if (wbase >= p.npaths) return;
  const uint32_t wcount = (uint32_t)min(per, p.npaths - wbase);
  const unsigned long long gbase = (unsigned long long)(p.path0 + wbase);
  const uint32_t hor = (uint32_t)min(p.horizon, 0xfffffffLL);
  const int na = p.t.na;
  uint32_t next = 0, q = 0, plo = 0, phi = 0, t = 0, a0 = 0;
  long long s = 0;
  float logw = 0.f;
  bool active = false;
  for (;;) {
    const unsigned need = __ballot_sync(FULL, !active);
    if (need) {
      const uint32_t rank = __popc(need & lt_mask);
      if (!active && next + rank < wcount) {
        q = next + rank;
        const unsigned long long gp = gbase + q;
        plo = (uint32_t)gp; phi = (uint32_t)(gp >> 32);
        s = p.starts ? p.starts[gp % (unsigned long long)p.nstarts] : (long long)(gp % (unsigned long long)p.nstarts);
        t = 0; a0 = 0; logw = 0.f;
        active = true;
      }
      next = min(next + (uint32_t)__popc(need), wcount);  // saturating: the counter must not wrap while the last lanes walk on
      if (!__any_sync(FULL, active)) break;
    }
    if (active) {
      bool done;
      if (p.t.term[s]) {
        logw = __fsub_rn(logw, p.vterm[s]);
        done = true;
      } else if (t >= hor) {
        done = true;
      } else {
        const uint4 r = philox4x32_10(p.keys, t, 0u, plo, phi);
        const float u0 = u01(r.x), u2 = u01(r.z);
        // StatsBase sample(a, Weights): i = 1; cw = w[1]; while cw < t && i < n: i += 1; cw += w[i]
        int a = 0;
        while (a < na - 1 && p.cdf[s * na + a] < u0) a++;
        logw = __fadd_rn(logw, p.bf[s * na + a]);
        if (t == 0u) a0 = (uint32_t)a;
        done = (r.y >> 8) >= p.gamma_k;  // u1 >= gamma (exact integer form)
        if (!done) {
          long long e = p.t.ptr[s * na + a];
          const long long last = p.t.ptr[s * na + a + 1] - 1;
          while (e < last && p.cum[e] < u2) e++;
          s = p.t.succ[e];
          t++;
        }
      }
      if (done) {
        p.logw[wbase + q] = logw;
        p.meta[wbase + q] = a0 | (t << 4);
        active = false;
      }
    }
  }
}

// ---- reduction into (max, sum exp(x - max), count) per (start slot, first action): two passes with atomics, buckets kept
// in shared memory per block when they fit (the single-start-state case would otherwise hammer a handful of addresses).
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
constexpr int TAB_SMEM_BUCKETS = 2048;

__global__ void table_reduce_init_kernel(int* omax, double* sum, unsigned long long* cnt, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { omax[i] = f2ord(-CUDART_INF_F); sum[i] = 0.0; cnt[i] = 0ull; }
}
__global__ void __launch_bounds__(256)
table_reduce_max_kernel(const float* __restrict__ logw, const uint32_t* __restrict__ meta, long long path0, long long npaths,
                        long long nstarts, int na, int* __restrict__ omax, unsigned long long* __restrict__ transitions) {
  __shared__ int sm[TAB_SMEM_BUCKETS];
  __shared__ unsigned long long red[32];
  const long long nb = nstarts * na;
  const bool local = nb <= TAB_SMEM_BUCKETS;
  if (local) { for (int i = threadIdx.x; i < nb; i += blockDim.x) sm[i] = f2ord(-CUDART_INF_F); __syncthreads(); }
  unsigned long long tr = 0;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npaths; q += (long long)gridDim.x * blockDim.x) {
    const uint32_t me = meta[q];
    tr += me >> 4;
    const long long b = ((path0 + q) % nstarts) * na + (me & 15u);
    if (local) atomicMax(&sm[b], f2ord(logw[q])); else atomicMax(&omax[b], f2ord(logw[q]));
  }
  if (local) { __syncthreads(); for (int i = threadIdx.x; i < nb; i += blockDim.x) atomicMax(&omax[i], sm[i]); }
  const unsigned long long r = block_sum(tr, red);
  if (threadIdx.x == 0 && r) atomicAdd(transitions, r);
}
__global__ void __launch_bounds__(256)
table_reduce_sum_kernel(const float* __restrict__ logw, const uint32_t* __restrict__ meta, long long path0, long long npaths,
                        long long nstarts, int na, const int* __restrict__ omax, double* __restrict__ sum,
                        unsigned long long* __restrict__ cnt) {
  __shared__ double ss[TAB_SMEM_BUCKETS];
  __shared__ unsigned int sc[TAB_SMEM_BUCKETS];
  const long long nb = nstarts * na;
  const bool local = nb <= TAB_SMEM_BUCKETS;
  if (local) { for (int i = threadIdx.x; i < nb; i += blockDim.x) { ss[i] = 0.0; sc[i] = 0u; } __syncthreads(); }
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npaths; q += (long long)gridDim.x * blockDim.x) {
    const long long b = ((path0 + q) % nstarts) * na + (meta[q] & 15u);
    const float m = ord2f(omax[b]), x = logw[q];
    const double e = x == -CUDART_INF_F ? 0.0 : (double)expf(x - m);
    if (local) { atomicAdd(&ss[b], e); atomicAdd(&sc[b], 1u); } else { atomicAdd(&sum[b], e); atomicAdd(&cnt[b], 1ull); }
  }
  if (local) {
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
      if (sc[i]) { atomicAdd(&sum[i], ss[i]); atomicAdd(&cnt[i], (unsigned long long)sc[i]); }
  }
}
// buckets -> per-(state, action) and per-state log mean weights
__global__ void table_finalize_kernel(const int* __restrict__ omax, const double* __restrict__ sum, const unsigned long long* __restrict__ cnt,
                                      const int* __restrict__ starts, long long nstarts, int na, double* __restrict__ logz_sa,
                                      unsigned long long* __restrict__ count_sa, double* __restrict__ logz_s) {
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nstarts) return;
  const long long st = starts ? starts[slot] : slot;
  double mm = -CUDART_INF;
  unsigned long long ctot = 0;
  for (int a = 0; a < na; a++) {
    const long long i = slot * na + a;
    const double m = (double)ord2f(omax[i]);
    if (logz_sa) logz_sa[st * na + a] = cnt[i] ? m + log(sum[i]) - log((double)cnt[i]) : CUDART_NAN;
    if (count_sa) count_sa[st * na + a] = cnt[i];
    if (cnt[i] && m > mm) mm = m;
    ctot += cnt[i];
  }
  if (logz_s) {
    double tot = 0.0;
    for (int a = 0; a < na; a++) {
      const long long i = slot * na + a;
      if (cnt[i]) tot += sum[i] * exp((double)ord2f(omax[i]) - mm);
    }
    logz_s[st] = ctot ? mm + log(tot) - log((double)ctot) : CUDART_NAN;
  }
}

void launch_table_policy(const TableDev& t, double scale, double beta, const double* V, const double* Q, float* cdf, float* bf,
                         float* cum, float* vterm, cudaStream_t st) {
  table_policy_kernel<<<tblk(t.ns), TAB_THREADS, 0, st>>>(t, scale, beta, V, Q, cdf, bf, vterm);
  table_cum_kernel<<<tblk(t.ns * t.na), TAB_THREADS, 0, st>>>(t, cum);
  g_launches += 2;
}
void launch_table_rollout(const TableDev& t, const float* cdf, const float* bf, const float* cum, const float* vterm, double gamma,
                          long long horizon, long long path0, long long npaths, const int* starts, long long nstarts,
                          uint64_t seed, float* logw, uint32_t* meta, int num_sms, cudaStream_t st) {
  if (npaths <= 0) return;
  TableRollArgs a{};
  a.t = t; a.cdf = cdf; a.bf = bf; a.cum = cum; a.vterm = vterm;
  a.gamma_k = (uint32_t)ceil(gamma * 16777216.0);  // u1 >= gamma against the Float64 gamma, for u1 = k 2^-24
  a.horizon = horizon; a.path0 = path0; a.npaths = npaths; a.starts = starts; a.nstarts = nstarts;
  a.keys = philox_expand((uint32_t)seed, (uint32_t)(seed >> 32));
  a.logw = logw; a.meta = meta;
  const long long want = (npaths + 255) / 256, cap = 8LL * num_sms;
  table_rollout_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a);
  g_launches++;
}
void launch_table_reduce(const float* logw, const uint32_t* meta, long long path0, long long npaths, const int* starts,
                         long long nstarts, int na, int* omax, double* sum, unsigned long long* cnt,
                         unsigned long long* transitions, double* logz_sa, unsigned long long* count_sa, double* logz_s,
                         int num_sms, cudaStream_t st) {
  const long long nb = nstarts * na;
  table_reduce_init_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(omax, sum, cnt, nb);
  const long long want = (npaths + 255) / 256, cap = 4LL * num_sms;
  const unsigned grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
  table_reduce_max_kernel<<<grid, 256, 0, st>>>(logw, meta, path0, npaths, nstarts, na, omax, transitions);
  table_reduce_sum_kernel<<<grid, 256, 0, st>>>(logw, meta, path0, npaths, nstarts, na, omax, sum, cnt);
  table_finalize_kernel<<<(unsigned)((nstarts + 255) / 256), 256, 0, st>>>(omax, sum, cnt, starts, nstarts, na, logz_sa, count_sa, logz_s);
  g_launches += 4;
}

}  // namespace piql
