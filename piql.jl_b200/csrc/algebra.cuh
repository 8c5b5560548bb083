// algebra.cuh -- the per-cell KL-control algebra shared by the backup kernels (sm_100a only).
//   truevalue_recursion src/PIQL.jl:95-101, logz_recursion :125-130, controlV :63-66,
//   optimality_gap_recursion src/value_iteration.jl:160-169, pa / log_pa / reward src/Problems/GridWorld.jl:98-116
#pragma once
#include <cmath>

#include "common.cuh"

namespace piql {

// Constants of one sweep in the sweep's precision.
template <typename T> struct K {
  int H, n, goal;
  T w_on, w_off;        // (a==aa)*(1-randomness) + randomness/len      GridWorld.jl:101
  T lw_on, lw_off;      // their logs (log_pa, GridWorld.jl:106)
  T gamma, lgamma, l1mgamma;
  T rwd_goal, rwd_step; // reward_function values times the modify() scale   GridWorld.jl:116
  T log_prior;          // log(1/len)                                       GridWorld.jl:112
  T beta;
};
template <typename T> inline K<T> make_k(const GridDev& g, double reward_scale, double beta) {
  K<T> k;
  k.H = g.H; k.n = g.n; k.goal = g.goal;
  const double w_on = 1.0 * (1.0 - g.randomness) + g.randomness / NA;
  const double w_off = 0.0 * (1.0 - g.randomness) + g.randomness / NA;
  k.w_on = (T)w_on; k.w_off = (T)w_off;
  k.lw_on = (T)log(w_on); k.lw_off = (T)log(w_off);
  k.gamma = (T)g.gamma; k.lgamma = (T)log(g.gamma); k.l1mgamma = (T)log(1.0 - g.gamma);
  k.rwd_goal = (T)(reward_scale * g.reward); k.rwd_step = (T)(reward_scale * -g.step_cost);
  k.log_prior = (T)log(1.0 / NA);
  k.beta = (T)beta;
  return k;
}

// ---- per-cell algebra on a gathered 5-point stencil ----------------------------------------------
template <typename T>
__device__ __forceinline__ void gather(const T* __restrict__ X, int s, int H, unsigned ci, T& c, T (&nb)[NA]) {
  // f(take_step(s,aa)): the neighbour, or the cell itself when a wall blocks the move  GridWorld.jl:29-32
  c = X[s];
  nb[0] = (ci & 1u) ? c : X[s + 1];
  nb[1] = (ci & 2u) ? c : X[s - 1];
  nb[2] = (ci & 4u) ? c : X[s + H];
  nb[3] = (ci & 8u) ? c : X[s - H];
}
template <typename T> __device__ __forceinline__ T pa(const K<T>& k, int a, const T (&f)[NA]) {
  T out = 0;  // propagator_average, same accumulation order as GridWorld.jl:98-104
#pragma unroll
  for (int aa = 0; aa < NA; aa++) out += (a == aa ? k.w_on : k.w_off) * f[aa];
  return out;
}
template <typename T>
__device__ __forceinline__ void avg_rewards(const K<T>& k, int s, unsigned ci, T (&rbar)[NA]) {
  // average_reward(ctrl,s,a) = pa(s,a, s1 -> reward_function(s,a,s1))   PIQL.jl:55-57
  T r[NA];
  r[0] = (!(ci & 1u) && s + 1 == k.goal) ? k.rwd_goal : k.rwd_step;
  r[1] = (!(ci & 2u) && s - 1 == k.goal) ? k.rwd_goal : k.rwd_step;
  r[2] = (!(ci & 4u) && s + k.H == k.goal) ? k.rwd_goal : k.rwd_step;
  r[3] = (!(ci & 8u) && s - k.H == k.goal) ? k.rwd_goal : k.rwd_step;
#pragma unroll
  for (int a = 0; a < NA; a++) rbar[a] = pa(k, a, r);
}
template <typename T>
__device__ __forceinline__ void control_v(const K<T>& k, T c, const T (&nb)[NA], const T (&rbar)[NA], T (&cv)[NA]) {
#pragma unroll
  for (int a = 0; a < NA; a++) cv[a] = k.gamma * pa(k, a, nb) - c + rbar[a];  // PIQL.jl:63-66
}

template <typename T>
__device__ __forceinline__ T backup_nu(const K<T>& k, int s, unsigned ci, const T (&nb)[NA]) {
  T rbar[NA], x[NA];
  avg_rewards(k, s, ci, rbar);
#pragma unroll
  for (int a = 0; a < NA; a++) x[a] = k.log_prior + k.gamma * pa(k, a, nb) + rbar[a];
  return lse4(x);
}
// logz_recursion (PIQL.jl:125-130) evaluated on Z = exp(logz): with A_a = log prior + beta*controlV + (1-beta)*controlQ
// (fixed during the solve; prep_kernel stores ea_a = exp(A_a - amax) and amax = max_a A_a),
//   Z_new = sum_a exp(A_a) * (gamma * sum_a' P(a'|a) Z(n_a') + 1 - gamma),
// with the neighbour maximum m and M = max(m, 0) factored out so that every exponent is <= 0.  Eight transcendental calls
// per cell with the damping, against 23 for the term-by-term logsumexp / logaddexp form of the same recursion.
template <typename T>
__device__ __forceinline__ T backup_logz(const K<T>& k, const T (&ea)[NA], T amax, const T (&nb)[NA]) {
  const T m = fmax(fmax(nb[0], nb[1]), fmax(nb[2], nb[3]));
  T e[NA], esum = 0;
#pragma unroll
  for (int aa = 0; aa < NA; aa++) { e[aa] = xexp(nb[aa] - m); esum += e[aa]; }
  const T em = xexp(-fabs(m));
  const bool pos = m >= (T)0;
  const T g1 = k.gamma * (pos ? (T)1 : em), g0 = ((T)1 - k.gamma) * (pos ? em : (T)1);
  T sum = 0;
#pragma unroll
  for (int a = 0; a < NA; a++) sum += ea[a] * (g1 * (k.w_on * e[a] + k.w_off * (esum - e[a])) + g0);
  return fmax(m, (T)0) + amax + xlog(sum);
}
// z_recursion (PIQL.jl:109-115), the linear-space form the reference keeps beside logz_recursion ("not numerically
// stable enough" in its own words): Z_new = sum_a exp(A_a) * (gamma * sum_a' P(a'|a) Z(n_a') + 1 - gamma), with
// eA_a = exp(A_a) stored by prep_kernel.
template <typename T>
__device__ __forceinline__ T backup_z(const K<T>& k, const T (&eA)[NA], const T (&nb)[NA]) {
  T out = 0;
#pragma unroll
  for (int a = 0; a < NA; a++) out += eA[a] * (k.gamma * pa(k, a, nb) + (T)1 - k.gamma);
  return out;
}
template <typename T>
__device__ __forceinline__ T backup_gap(const K<T>& k, const T (&P)[NA], T D, const T (&nb)[NA]) {
  T out = D;  // D = sum_a pi_a (controlQ - controlV(nu)); gap term has no gamma (value_iteration.jl:165)
#pragma unroll
  for (int a = 0; a < NA; a++) out += P[a] * pa(k, a, nb);
  return out;
}

}  // namespace piql
