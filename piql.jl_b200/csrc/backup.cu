// backup.cu -- soft value-iteration Bellman backups on the implicit gridworld (sm_100a).
//
// Restates, per cell, the recursions the reference evaluates through closures and a Dict:
//   nu    truevalue_recursion        src/PIQL.jl:95-101      driven by update_ν!   src/value_iteration.jl:62-75
//   logz  logz_recursion             src/PIQL.jl:125-130     driven by update_logz! src/value_iteration.jl:89-105
//   gap   optimality_gap_recursion   src/value_iteration.jl:160-169, update_optimality_gap! :129-142
// with the gridworld closures pa / log_pa / reward (src/Problems/GridWorld.jl:98-116).
//
// Layout: one value per cell of the H x W wall array (border included), column-major, so a
// thread per cell reads its own value and the 4 neighbours at s+-1, s+-H: loads are coalesced
// along i and the +-H neighbours are whole coalesced lines served from L1/L2.  Sweeps are Jacobi
// (ping-pong): every cell's new value depends only on the previous iterate, so the result is
// independent of grid/block shape and of how many GPUs a caller splits the grid over.
#include "internal.h"
#include "algebra.cuh"

namespace piql {

thread_local long long g_launches = 0;

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };
template <typename T> __device__ __forceinline__ void load4(const T* p, int s, T (&o)[NA]);
template <> __device__ __forceinline__ void load4<float>(const float* p, int s, float (&o)[NA]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p) + s);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load4<double>(const double* p, int s, double (&o)[NA]) {
  const double2 v0 = __ldg(reinterpret_cast<const double2*>(p) + 2 * s);
  const double2 v1 = __ldg(reinterpret_cast<const double2*>(p) + 2 * s + 1);
  o[0] = v0.x; o[1] = v0.y; o[2] = v1.x; o[3] = v1.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, int s, const T (&v)[NA]);
template <> __device__ __forceinline__ void store4<float>(float* p, int s, const float (&v)[NA]) {
  reinterpret_cast<float4*>(p)[s] = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<double>(double* p, int s, const double (&v)[NA]) {
  reinterpret_cast<double2*>(p)[2 * s] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[2 * s + 1] = make_double2(v[2], v[3]);
}

// One backup of cell s reading the iterate through X (global or shared memory).
template <typename T, int KIND>
__device__ __forceinline__ void backup_cell(const K<T>& k, int s, unsigned ci, const T* X, const T* __restrict__ aux4,
                                            const T* __restrict__ aux1, T term_value, T& stored, double& absdiff) {
  T c, nb[NA];
  gather(X, s, k.H, ci, c, nb);
  T nw;
  if (ci & CELL_GOAL) {
    nw = term_value;  // terminal_condition: nu, gap -> 0; logz -> 0 - V(goal); Z -> exp(0 - V(goal))
  } else if (KIND == KIND_NU) {
    nw = backup_nu(k, s, ci, nb);
  } else if (KIND == KIND_LOGZ) {
    T ea[NA];
    load4(aux4, s, ea);
    nw = backup_logz(k, ea, __ldg(aux1 + s), nb);
  } else if (KIND == KIND_Z) {
    T eA[NA];
    load4(aux4, s, eA);
    nw = backup_z(k, eA, nb);
  } else {
    T P[NA];
    load4(aux4, s, P);
    nw = backup_gap(k, P, __ldg(aux1 + s), nb);
  }
  if (KIND == KIND_Z) {  // update_z!: diff += |log(new) - log(Z(s0))|, Z <- eta*new + (1-eta)*Z   value_iteration.jl:35-37
    absdiff = fabs(log((double)nw) - log((double)c));
    stored = (T)0.5 * nw + ((T)1 - (T)0.5) * c;
    return;
  }
  absdiff = fabs((double)nw - (double)c);
  // eta = 0.5 damping of update_logz!: logaddexp(log(eta)+new, log(1-eta)+old)   value_iteration.jl:101-102
  stored = (KIND == KIND_LOGZ) ? logaddexp((T)(-0.6931471805599453) + nw, (T)(-0.6931471805599453) + c) : nw;
}

// ---- kernels -----------------------------------------------------------------------------------
__global__ void cellinfo_kernel(const uint8_t* __restrict__ walls, int H, int W, int goal, uint8_t* __restrict__ cell) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= H * W) return;
  const int i = s % H, j = s / H;
  unsigned ci = 0;
  if (walls[s]) {
    ci = CELL_WALL | CELL_BLOCK_MASK;
  } else {
    // the border is wall by construction (make_walls, GridWorld.jl:40-45); guard anyway
    if (i + 1 >= H || walls[s + 1]) ci |= 1u;
    if (i - 1 < 0 || walls[s - 1]) ci |= 2u;
    if (j + 1 >= W || walls[s + H]) ci |= 4u;
    if (j - 1 < 0 || walls[s - H]) ci |= 8u;
    if (s == goal) ci |= CELL_GOAL;
  }
  cell[s] = (uint8_t)ci;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(SWEEP_THREADS)
sweep_kernel(K<T> k, const uint8_t* __restrict__ cell, const T* __restrict__ X, T* __restrict__ Y,
             const T* __restrict__ aux4, const T* __restrict__ aux1, T term_value, double* __restrict__ partial) {
  __shared__ double red[32];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  if (s < k.n) {
    const unsigned ci = cell[s];
    if (!(ci & CELL_WALL)) {
      T stored;
      backup_cell<T, KIND>(k, s, ci, X, aux4, aux1, term_value, stored, d);
      Y[s] = stored;
    }
  }
  if (partial) {  // uniform branch
    const double r = block_sum(d, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
  }
}

__global__ void __launch_bounds__(1024) residual_final_kernel(const double* __restrict__ partial, int nblocks, double* out) {
  __shared__ double red[32];
  double d = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) d += partial[b];
  const double r = block_sum(d, red);
  if (threadIdx.x == 0) *out = r;
}

// Whole solve in one CTA with the iterate resident in shared memory: the reference-scale
// problems (251 states) need 10^3..10^4 sweeps, which one launch per sweep would turn into
// pure launch latency.
template <typename T, int KIND>
__global__ void __launch_bounds__(1024)
resident_solve_kernel(K<T> k, const uint8_t* __restrict__ cell, const T* __restrict__ x0, T* __restrict__ xout,
                      const T* __restrict__ aux4, const T* __restrict__ aux1, T term_value, double tol,
                      long long max_sweeps, double* __restrict__ out2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ double s_res;
  T* A = reinterpret_cast<T*>(smem_raw);
  T* B = A + k.n;
  uint8_t* cs = reinterpret_cast<uint8_t*>(B + k.n);
  for (int s = threadIdx.x; s < k.n; s += blockDim.x) { A[s] = x0[s]; B[s] = x0[s]; cs[s] = cell[s]; }
  __syncthreads();
  long long sweeps = 0;
  double res = CUDART_INF;
  while (sweeps < max_sweeps) {
    double d = 0.0;
    for (int s = threadIdx.x; s < k.n; s += blockDim.x) {
      const unsigned ci = cs[s];
      if (ci & CELL_WALL) continue;
      T stored; double dd;
      backup_cell<T, KIND>(k, s, ci, A, aux4, aux1, term_value, stored, dd);
      B[s] = stored;
      d += dd;
    }
    const double r = block_sum(d, red);  // contains the barriers that order A reads / B writes
    if (threadIdx.x == 0) s_res = r;
    __syncthreads();
    res = s_res;
    T* t = A; A = B; B = t;
    sweeps++;
    if (res < tol) break;
  }
  for (int s = threadIdx.x; s < k.n; s += blockDim.x) xout[s] = A[s];
  if (threadIdx.x == 0) { out2[0] = (double)sweeps; out2[1] = res; }
}


// ---- batched resident solves: one CTA per problem, everything in shared memory -----------------------------
// The experiment protocol of the reference (run_experiment, scripts/experiments.jl:20-39; learning_traj :10-18)
// is 8 independent beta chains of 5 dependent logz_updateV solves and then 48 independent optimality-gap solves
// on one 251-state problem: 10^3..10^4 sweeps each, far too small for more than one CTA but perfectly parallel
// across problems.  Each CTA keeps V, the iterate (ping-pong) and the per-cell constants in shared memory, builds
// the constants itself (what prep_kernel does for a single solve), iterates to the reference's stop rule, and --
// for a chain -- applies V <- V + logz/beta (logz_updateV, value_iteration.jl:119-124) and goes on.
// kind KIND_LOGZ: chain of nupdates solves, out[b][0..nupdates][n];  KIND_GAP / KIND_NU: one solve, out[b][n].
template <typename T>
__global__ void __launch_bounds__(1024)
resident_batch_kernel(K<T> k, int kind, const uint8_t* __restrict__ cell, const double* __restrict__ Vin, long long vin_stride,
                      const double* __restrict__ nu, const double* __restrict__ betas, int nupdates, double tol,
                      long long max_sweeps, double* __restrict__ out, double* __restrict__ stats /*[b][nupdates][2]*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ double s_res;
  const int n = k.n, b = blockIdx.x;
  T* V = reinterpret_cast<T*>(smem_raw);
  T* A = V + n;
  T* B = A + n;
  T* aux = B + n;  // five planes: logz A_a (4); gap pi_a (4), D
  uint8_t* cs = reinterpret_cast<uint8_t*>(aux + 5 * (size_t)n);
  if (betas) k.beta = (T)betas[b];
  const double beta_d = betas ? betas[b] : 1.0;
  for (int s = threadIdx.x; s < n; s += blockDim.x) {
    const double v = Vin ? Vin[(size_t)b * vin_stride + s] : 0.0;
    V[s] = (v == v) ? (T)v : (T)0;
    cs[s] = cell[s];
  }
  __syncthreads();
  const int nsolves = kind == KIND_LOGZ ? nupdates : 1;
  double* outb = out + (size_t)b * (kind == KIND_LOGZ ? (size_t)(nupdates + 1) * n : (size_t)n);
  if (kind == KIND_LOGZ)
    for (int s = threadIdx.x; s < n; s += blockDim.x) outb[s] = (cs[s] & CELL_WALL) ? CUDART_NAN : (double)V[s];
  for (int u = 0; u < nsolves; u++) {
    // per-cell constants from the current V (PIQL.jl:63-85)
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
      const unsigned ci = cs[s];
      T o4[NA] = {0, 0, 0, 0}, o1 = 0;
      if (kind != KIND_NU && !(ci & (CELL_WALL | CELL_GOAL))) {
        T c, nb[NA], rbar[NA], cv[NA], x[NA];
        gather(V, s, k.H, ci, c, nb);
        avg_rewards(k, s, ci, rbar);
        control_v(k, c, nb, rbar, cv);
#pragma unroll
        for (int a = 0; a < NA; a++) x[a] = k.log_prior + cv[a];
        const T norm = lse4(x);  // Qnormalization of the V-determined Q
        if (kind == KIND_LOGZ) {
#pragma unroll
          for (int a = 0; a < NA; a++) o4[a] = k.log_prior + k.beta * cv[a] + ((T)1 - k.beta) * (cv[a] - norm);
          o1 = fmax(fmax(o4[0], o4[1]), fmax(o4[2], o4[3]));
#pragma unroll
          for (int a = 0; a < NA; a++) o4[a] = xexp(o4[a] - o1);
        } else {
          T nc, nnb[NA], cvn[NA];
          nc = (T)nu[s];
          nnb[0] = (ci & 1u) ? nc : (T)nu[s + 1];
          nnb[1] = (ci & 2u) ? nc : (T)nu[s - 1];
          nnb[2] = (ci & 4u) ? nc : (T)nu[s + k.H];
          nnb[3] = (ci & 8u) ? nc : (T)nu[s - k.H];
          control_v(k, nc, nnb, rbar, cvn);
#pragma unroll
          for (int a = 0; a < NA; a++) {
            const T cq = cv[a] - norm;
            o4[a] = (T)0.25 * xexp(cq);
            o1 += o4[a] * (cq - cvn[a]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < NA; a++) aux[(size_t)a * n + s] = o4[a];
      aux[(size_t)4 * n + s] = o1;
      A[s] = (T)0;  // Growable default x -> 0.0
      B[s] = (T)0;
    }
    __syncthreads();
    const T term = kind == KIND_LOGZ ? (T)0 - V[k.goal] : (T)0;
    T* X = A;
    T* Y = B;
    long long sweeps = 0;
    double res = CUDART_INF;
    while (sweeps < max_sweeps) {
      double d = 0.0;
      for (int s = threadIdx.x; s < n; s += blockDim.x) {
        const unsigned ci = cs[s];
        if (ci & CELL_WALL) continue;
        T c, nb[NA], nw;
        gather(X, s, k.H, ci, c, nb);
        if (ci & CELL_GOAL) nw = term;
        else if (kind == KIND_NU) nw = backup_nu(k, s, ci, nb);
        else {
          const T a4[NA] = {aux[s], aux[(size_t)n + s], aux[(size_t)2 * n + s], aux[(size_t)3 * n + s]};
          nw = kind == KIND_LOGZ ? backup_logz(k, a4, aux[(size_t)4 * n + s], nb) : backup_gap(k, a4, aux[(size_t)4 * n + s], nb);
        }
        d += fabs((double)nw - (double)c);
        Y[s] = kind == KIND_LOGZ ? logaddexp((T)(-0.6931471805599453) + nw, (T)(-0.6931471805599453) + c) : nw;
      }
      const double r = block_sum(d, red);
      if (threadIdx.x == 0) s_res = r;
      __syncthreads();
      res = s_res;
      T* t = X; X = Y; Y = t;
      sweeps++;
      if (res < tol) break;
    }
    if (threadIdx.x == 0) { stats[((size_t)b * nsolves + u) * 2] = (double)sweeps; stats[((size_t)b * nsolves + u) * 2 + 1] = res; }
    if (kind == KIND_LOGZ) {
      double* o = outb + (size_t)(u + 1) * n;
      for (int s = threadIdx.x; s < n; s += blockDim.x) {
        const bool wall = cs[s] & CELL_WALL;
        // V(k) + logz(k)/beta in double from the stored iterate, like piql_update_v
        const double vn = (double)V[s] + (double)X[s] / beta_d;
        if (!wall) V[s] = (T)vn;
        o[s] = wall ? CUDART_NAN : (sizeof(T) == 8 ? vn : (double)V[s]);
      }
    } else {
      for (int s = threadIdx.x; s < n; s += blockDim.x) outb[s] = (cs[s] & CELL_WALL) ? CUDART_NAN : (double)X[s];
    }
    __syncthreads();
  }
}

size_t resident_batch_smem_bytes(int n, int precision) {
  const size_t e = precision == 0 ? 8 : 4;
  return 8 * e * (size_t)n + (((size_t)n + 15) / 16) * 16;
}
void launch_resident_batch(const GridDev& g, const uint8_t* cell, int kind, int precision, double reward_scale, int nproblems,
                           const double* Vin, long long vin_stride, const double* nu, const double* betas, int nupdates,
                           double tol, long long max_sweeps, double* out, double* stats, cudaStream_t st) {
  const size_t smem = resident_batch_smem_bytes(g.n, precision);
  const int threads = g.n >= 1024 ? 1024 : ((g.n + 31) / 32) * 32;
  if (precision == 0)
    resident_batch_kernel<double><<<nproblems, threads, smem, st>>>(make_k<double>(g, reward_scale, 1.0), kind, cell, Vin, vin_stride,
                                                                    nu, betas, nupdates, tol, max_sweeps, out, stats);
  else
    resident_batch_kernel<float><<<nproblems, threads, smem, st>>>(make_k<float>(g, reward_scale, 1.0), kind, cell, Vin, vin_stride,
                                                                   nu, betas, nupdates, tol, max_sweeps, out, stats);
  g_launches++;
}

// Per-cell constants of the logz / gap sweeps.
// Always evaluated in double from the caller's Float64 V / Q / nu and rounded ONCE to the sweep precision: the constants are
// per-cell inputs of every sweep, so a rounding pattern shared by all cells (fp32 log 1/4, fp32 gamma ...) would act as a
// bias the recursion amplifies by 1/(1-gamma) -- measured 2e-5 relative on log Z at beta = 0.2 with fp32 constants.
template <typename TO, int KIND>
__global__ void __launch_bounds__(SWEEP_THREADS)
prep_kernel(K<double> k, const uint8_t* __restrict__ cell, const double* __restrict__ V, const double* __restrict__ Q,
            const double* __restrict__ nu, TO* __restrict__ aux4o, TO* __restrict__ aux1o) {
  typedef double T;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= k.n) return;
  const unsigned ci = cell[s];
  T out4[NA] = {0, 0, 0, 0};
  T out1 = 0;
  if (!(ci & (CELL_WALL | CELL_GOAL))) {
    T c, nb[NA], rbar[NA], cv[NA], q[NA], x[NA];
    gather(V, s, k.H, ci, c, nb);
    avg_rewards(k, s, ci, rbar);
    control_v(k, c, nb, rbar, cv);
    if (Q) load4(Q, s, q);
    else {
#pragma unroll
      for (int a = 0; a < NA; a++) q[a] = cv[a];  // (s,a) -> controlV(ctrl,V,s,a)   value_iteration.jl:97,157
    }
#pragma unroll
    for (int a = 0; a < NA; a++) x[a] = k.log_prior + q[a];
    const T norm = lse4(x);  // Qnormalization   PIQL.jl:72-75
    if (KIND == KIND_Z) {  // prior * exp(beta*controlV + (1-beta)*controlQ)   PIQL.jl:112
#pragma unroll
      for (int a = 0; a < NA; a++) out4[a] = (T)0.25 * xexp(k.beta * cv[a] + ((T)1 - k.beta) * (q[a] - norm));
    } else if (KIND == KIND_LOGZ) {
#pragma unroll
      for (int a = 0; a < NA; a++) out4[a] = k.log_prior + k.beta * cv[a] + ((T)1 - k.beta) * (q[a] - norm);
      out1 = fmax(fmax(out4[0], out4[1]), fmax(out4[2], out4[3]));   // stored as exp(A_a - max A) and max A
#pragma unroll
      for (int a = 0; a < NA; a++) out4[a] = xexp(out4[a] - out1);
    } else {
      T nc, nnb[NA], cvn[NA];
      gather(nu, s, k.H, ci, nc, nnb);
      control_v(k, nc, nnb, rbar, cvn);
#pragma unroll
      for (int a = 0; a < NA; a++) {
        const T cq = q[a] - norm;                 // controlQ            PIQL.jl:79-81
        out4[a] = (T)0.25 * xexp(cq);             // action_probability  PIQL.jl:83-85
        out1 += out4[a] * (cq - cvn[a]);
      }
    }
  }
  TO o4[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) o4[a] = (TO)out4[a];
  store4(aux4o, s, o4);
  aux1o[s] = (TO)out1;
}

template <typename T>
__global__ void convert_in_kernel(const double* __restrict__ src, T* __restrict__ dst, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) {
    const double v = src[s];
    dst[s] = (v == v) ? (T)v : (T)0;  // callers pass NaN at walls; never read, but keep them finite
  }
}
template <typename T>
__global__ void convert_out_kernel(const T* __restrict__ src, const uint8_t* __restrict__ cell, double* __restrict__ dst, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[s] = (cell[s] & CELL_WALL) ? CUDART_NAN : (double)src[s];
}
template <typename T> __global__ void fill_kernel(T* dst, T v, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) dst[s] = v;
}
template <typename T>
__global__ void update_v_kernel(const double* __restrict__ V, const T* __restrict__ logz, const uint8_t* __restrict__ cell,
                                double beta, double* __restrict__ out, int n, int linear) {
  // V(k) + logz(k)/beta   value_iteration.jl:121-122;  linear: V(k) + log(Z(k))/beta   :55-57
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) out[s] = (cell[s] & CELL_WALL) ? CUDART_NAN : V[s] + (linear ? log((double)logz[s]) : (double)logz[s]) / beta;
}

// ---- host launchers ------------------------------------------------------------------------------
static inline int nblk(int n, int t) { return (n + t - 1) / t; }

void launch_cellinfo(const uint8_t* walls, const GridDev& g, uint8_t* cell, cudaStream_t st) {
  cellinfo_kernel<<<nblk(g.n, 256), 256, 0, st>>>(walls, g.H, g.W, g.goal, cell);
  g_launches++;
}
void launch_convert_in(const double* src, void* dst, int n, int precision, cudaStream_t st) {
  if (precision == 0) convert_in_kernel<double><<<nblk(n, 256), 256, 0, st>>>(src, (double*)dst, n);
  else convert_in_kernel<float><<<nblk(n, 256), 256, 0, st>>>(src, (float*)dst, n);
  g_launches++;
}
void launch_convert_out(const void* src, const uint8_t* cell, double* dst, int n, int precision, cudaStream_t st) {
  if (precision == 0) convert_out_kernel<double><<<nblk(n, 256), 256, 0, st>>>((const double*)src, cell, dst, n);
  else convert_out_kernel<float><<<nblk(n, 256), 256, 0, st>>>((const float*)src, cell, dst, n);
  g_launches++;
}
void launch_fill(void* dst, double value, int n, int precision, cudaStream_t st) {
  if (precision == 0) fill_kernel<double><<<nblk(n, 256), 256, 0, st>>>((double*)dst, value, n);
  else fill_kernel<float><<<nblk(n, 256), 256, 0, st>>>((float*)dst, (float)value, n);
  g_launches++;
}
void launch_update_v(const double* V, const void* logz, const uint8_t* cell, double beta, double* out, int n,
                     int precision, cudaStream_t st, bool linear) {
  if (precision == 0) update_v_kernel<double><<<nblk(n, 256), 256, 0, st>>>(V, (const double*)logz, cell, beta, out, n, linear);
  else update_v_kernel<float><<<nblk(n, 256), 256, 0, st>>>(V, (const float*)logz, cell, beta, out, n, linear);
  g_launches++;
}

template <typename T>
static void prep_t(const GridDev& g, const uint8_t* cell, ViState& vs, const double* V, const double* Q, const double* nu,
                   cudaStream_t st) {
  const K<double> k = make_k<double>(g, vs.reward_scale, vs.beta);
  const int nb = nblk(g.n, SWEEP_THREADS);
  if (vs.kind == KIND_LOGZ)
    prep_kernel<T, KIND_LOGZ><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, V, Q, nullptr, (T*)vs.aux4, (T*)vs.aux1);
  else if (vs.kind == KIND_Z)
    prep_kernel<T, KIND_Z><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, V, Q, nullptr, (T*)vs.aux4, (T*)vs.aux1);
  else
    prep_kernel<T, KIND_GAP><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, V, Q, nu, (T*)vs.aux4, (T*)vs.aux1);
  g_launches++;
}
// V, Q (nullable), nu: the caller's Float64 arrays on the device (NaN at walls is fine: wall cells are never read)
void launch_prep(const GridDev& g, const uint8_t* cell, ViState& vs, const double* V, const double* Q, const double* nu,
                 cudaStream_t st) {
  if (vs.kind == KIND_NU) return;
  if (vs.precision == 0) prep_t<double>(g, cell, vs, V, Q, nu, st);
  else prep_t<float>(g, cell, vs, V, Q, nu, st);
}

// terminal cells: nu, gap -> 0; logz -> 0 - V(goal)   value_iteration.jl:66-67,93-94,133-134
template <typename T> static T term_value_of(const ViState& vs) {
  if (vs.kind == KIND_Z) return (T)exp(0.0 - vs.v_goal);  // value_iteration.jl:28
  return vs.kind == KIND_LOGZ ? (T)(0.0 - vs.v_goal) : (T)0;
}

template <typename T>
static void sweep_t(const GridDev& g, const uint8_t* cell, ViState& vs, bool with_residual, cudaStream_t st) {
  const K<T> k = make_k<T>(g, vs.reward_scale, vs.beta);
  const int nb = vs.nblocks;
  const T* X = (const T*)vs.x[vs.cur];
  T* Y = (T*)vs.x[vs.cur ^ 1];
  double* partial = with_residual ? vs.partial : nullptr;
  const T tv = term_value_of<T>(vs);
  switch (vs.kind) {
    case KIND_NU:
      sweep_kernel<T, KIND_NU><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, X, Y, nullptr, nullptr, tv, partial); break;
    case KIND_LOGZ:
      sweep_kernel<T, KIND_LOGZ><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, X, Y, (const T*)vs.aux4, (const T*)vs.aux1, tv, partial); break;
    case KIND_Z:
      sweep_kernel<T, KIND_Z><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, X, Y, (const T*)vs.aux4, (const T*)vs.aux1, tv, partial); break;
    default:
      sweep_kernel<T, KIND_GAP><<<nb, SWEEP_THREADS, 0, st>>>(k, cell, X, Y, (const T*)vs.aux4, (const T*)vs.aux1, tv, partial); break;
  }
  g_launches++;
  if (with_residual) {
    residual_final_kernel<<<1, 1024, 0, st>>>(vs.partial, nb, vs.residual);
    g_launches++;
  }
  vs.cur ^= 1;
}

void launch_residual_final(const double* partial, int nblocks, double* out, cudaStream_t st) {
  residual_final_kernel<<<1, 1024, 0, st>>>(partial, nblocks, out);
  g_launches++;
}

void launch_sweep(const GridDev& g, const uint8_t* cell, ViState& vs, bool with_residual, cudaStream_t st) {
  if (vs.precision == 0) sweep_t<double>(g, cell, vs, with_residual, st);
  else sweep_t<float>(g, cell, vs, with_residual, st);
}

size_t resident_smem_bytes(int n, int precision) {
  const size_t e = precision == 0 ? 8 : 4;
  return 2 * e * (size_t)n + (((size_t)n + 15) / 16) * 16;
}

template <typename T>
static void resident_t(const GridDev& g, const uint8_t* cell, ViState& vs, double tol, long long max_sweeps,
                       double* dev_out, cudaStream_t st) {
  const K<T> k = make_k<T>(g, vs.reward_scale, vs.beta);
  const size_t smem = resident_smem_bytes(g.n, vs.precision);
  const T* x0 = (const T*)vs.x[vs.cur];
  T* xout = (T*)vs.x[vs.cur ^ 1];
  const T tv = term_value_of<T>(vs);
  const int threads = g.n >= 1024 ? 1024 : ((g.n + 31) / 32) * 32;
  switch (vs.kind) {
    case KIND_NU:
      resident_solve_kernel<T, KIND_NU><<<1, threads, smem, st>>>(k, cell, x0, xout, nullptr, nullptr, tv, tol, max_sweeps, dev_out); break;
    case KIND_LOGZ:
      resident_solve_kernel<T, KIND_LOGZ><<<1, threads, smem, st>>>(k, cell, x0, xout, (const T*)vs.aux4, (const T*)vs.aux1, tv, tol, max_sweeps, dev_out); break;
    case KIND_Z:
      resident_solve_kernel<T, KIND_Z><<<1, threads, smem, st>>>(k, cell, x0, xout, (const T*)vs.aux4, (const T*)vs.aux1, tv, tol, max_sweeps, dev_out); break;
    default:
      resident_solve_kernel<T, KIND_GAP><<<1, threads, smem, st>>>(k, cell, x0, xout, (const T*)vs.aux4, (const T*)vs.aux1, tv, tol, max_sweeps, dev_out); break;
  }
  g_launches++;
  vs.cur ^= 1;
}
void launch_resident_solve(const GridDev& g, const uint8_t* cell, ViState& vs, double tol, long long max_sweeps,
                           double* dev_out, cudaStream_t st) {
  if (vs.precision == 0) resident_t<double>(g, cell, vs, tol, max_sweeps, dev_out, st);
  else resident_t<float>(g, cell, vs, tol, max_sweeps, dev_out, st);
}

int configure_backup_kernels() {
  // opt in to > 48 KB dynamic shared memory for the resident solver
  const int maxb = (int)resident_smem_bytes(RESIDENT_MAX_CELLS, 0);
  cudaError_t e = cudaSuccess;
#define PIQL_SET(T, KND) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(resident_solve_kernel<T, KND>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb);
  PIQL_SET(double, KIND_NU) PIQL_SET(double, KIND_LOGZ) PIQL_SET(double, KIND_GAP) PIQL_SET(double, KIND_Z)
  PIQL_SET(float, KIND_NU) PIQL_SET(float, KIND_LOGZ) PIQL_SET(float, KIND_GAP) PIQL_SET(float, KIND_Z)
#undef PIQL_SET
  if (e == cudaSuccess) e = cudaFuncSetAttribute(resident_batch_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(resident_batch_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return (int)e;
}

}  // namespace piql
