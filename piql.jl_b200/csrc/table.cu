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
                  double* __restrict__ aux, double* __restrict__ aux1) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= t.ns) return;
  const int na = t.na;
  double cv[TAB_MAX_ACTIONS], x[TAB_MAX_ACTIONS];
  for (int a = 0; a < na; a++) {
    cv[a] = t.gamma * tab_pa(t, s * na + a, V) - V[s] + tab_rbar(t, s * na + a, scale);  // controlV  PIQL.jl:63-66
    x[a] = t.lprior[s * na + a] + cv[a];
  }
  const double norm = tab_lse(x, na);  // Qnormalization  PIQL.jl:72-75
  double D = 0.0;
  for (int a = 0; a < na; a++) {
    const double cq = cv[a] - norm;  // controlQ  PIQL.jl:79-81
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
                       double* aux1, cudaStream_t st) {
  table_prep_kernel<<<tblk(t.ns), TAB_THREADS, 0, st>>>(t, kind, scale, beta, V, nu, aux, aux1);
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
