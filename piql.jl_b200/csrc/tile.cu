// tile.cu -- temporally blocked soft value-iteration sweeps for grids that do not fit one CTA (sm_100a).
//
// The same three recursions as backup.cu (truevalue_recursion src/PIQL.jl:95-101 driven by update_ν!
// src/value_iteration.jl:62-75; logz_recursion src/PIQL.jl:125-130 driven by update_logz! :89-105;
// optimality_gap_recursion :160-169 driven by update_optimality_gap! :129-142), restructured for the
// machine: a value array of 10^6..10^7 cells is 4..67 MB, L2-resident, and one Jacobi sweep at the HBM
// roofline lasts 1.3..20 us -- below a kernel launch.  So one launch performs TILE_KH sweeps:
//
//   * the grid lives in a padded layout (ghost frame of TILE_KH cells, leading dimension a multiple of
//     32 floats) so that every tile load is an aligned, unpredicated 128-bit access;
//   * a CTA stages a tile of 128 x EJ cells (interior 120 x TJ + halo TILE_KH) and its 1-byte wall
//     masks in shared memory, runs TILE_KH Jacobi sweeps there (ping-pong), the valid region shrinking
//     by one ring per sweep, and writes the interior back: global traffic per sweep drops by TILE_KH;
//   * a warp owns a column: lane l holds the 4-cell strip of rows 4l..4l+3 in registers (one LDS.128
//     for itself, one each for the left and right columns), the vertical neighbours across strips come
//     from two warp shuffles, and the wall test is a bit of the mask word -> a select;
//   * per-cell algebra is refactored so that the four exponentials are shared between the actions
//     (P(a'|a) takes two values) and, in the fast fp32 mode, map 1:1 onto MUFU.EX2 / MUFU.LG2.
//
// Jacobi makes every cell's update a pure function of the previous iterate, so the values are
// independent of the tiling (halo cells recomputed by two CTAs get identical bits).
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "algebra.cuh"
#include "dtables.cuh"
#include "internal.h"

namespace piql {

#ifndef PIQL_TILE_UNROLL
#define PIQL_TILE_UNROLL 1  // columns of a sub-sweep a warp interleaves; measured on B200: 1 beats 2, 3 and 4 by 3-6 % on every kind and precision
#endif
constexpr int TILE_UNROLL = PIQL_TILE_UNROLL;
// launch bounds of the tiled kernels: threads per CTA, CTAs per SM (registers per thread <= 65536 / their product)
#ifndef PIQL_TILE_F32_MAXT
#define PIQL_TILE_F32_MAXT 512
#endif
#ifndef PIQL_TILE_F32_MINB
#define PIQL_TILE_F32_MINB 2
#endif
#ifndef PIQL_TILE_F64_MAXT
#define PIQL_TILE_F64_MAXT 512
#endif
#ifndef PIQL_TILE_F64_MINB
#define PIQL_TILE_F64_MINB 1
#endif
constexpr int TEI = 128;            // rows of the shared-memory tile: 32 lanes x 4-cell strips
// Cell bytes of a staged column: 128 + 16.  A TMA box must start on a 16-byte boundary and a tile starts at byte
// 120*ti of the byte array, so the box starts up to 8 bytes early and the kernel indexes past that offset.
constexpr int TMB = TEI + 16;       // bytes
constexpr int TMP = TMB / 4;        // 32-bit words

// Constants of the tiled sweeps.
template <typename T> struct TileK {
  int ld, np;          // leading dimension and size of the padded arrays
  int TJ, EJ, tiles_i;
  int kh, tti;         // sweeps per round = ghost depth (4 or 8); interior rows per tile = 128 - 2 * kh
  int tj0;             // first tile column of this launch (multi-GPU column slabs; 0 otherwise)
  int goal_pi, goal_pj, goal_p;  // goal in padded coordinates
  T kc, a1, b1, a2;    // nu: x_a = kc + a1*n_a + b1*S;  a2 = a1*log2(e)
  T kc_lo, a1_lo;      // fp32: what rounding kc and a1 to float lost (a constant per-sweep bias is amplified by 1/(1-gamma))
  T w_off, dw;         // P(a'|a) = w_off + dw*[a'==a]
  T gamma, omg;        // gamma, 1 - gamma
  T term;              // value of the terminal cell
  K<T> k;              // constants of the general formula (H := ld, goal := goal_p), for the cells next to the goal
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- table-driven Float64 exp / log for the tiled sweeps --------------------------------------------------------------
// The Float64 sweeps are bound by the FP64 pipe, and two thirds of their DFMA are the library's exp / log (~21 each, plus the
// special-case branches).  Here the arguments are known: exponents are <= 0 (everything is shifted by a maximum), logarithms
// take sums of such terms.  exp: x = k ln2/256 + r, e^x = 2^(k >> 8) * T[k & 255] * P5(r), |r| <= ln2/512 (Cody-Waite
// reduction, degree-5 Taylor: truncation 8e-21).  log: s = 2^e m, m in [1,2) looked up in 128 intervals, u = m / c_i - 1,
// log s = e ln2 + log c_i + (u - u^2/2 + ... + u^7/7), |u| <= 1/256 (truncation 7e-21).  About 10 FP64 operations each and no
// branches at all; absolute error ~3e-16 (tables: tools/gen_dtables.py, 64-bit-mantissa arithmetic).  Special values (NaN,
// -inf exponent, zero / subnormal / infinite / negative sum) are patched in with selects so that they propagate as the
// library's would.
__device__ __forceinline__ double tile_dexp_np(double x) {
  // branch-free (a branch per call would fence the scheduler: the 16 calls of a strip are independent and must interleave).
  // Domain: x <= 0, -inf included, or NaN (propagated); e^-708 is the smallest normal result, below it terms are ~0 anyway.
  const double xc = fmax(x, -708.0);                    // also maps NaN to a finite value: the result is patched below
  const double magic = 6755399441055744.0;              // 1.5 * 2^52: the sum's low word is round-to-nearest(x * 256/ln2)
  const double t = fma(xc, DEXP_INVL, magic);
  const int k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(-kd, DEXP_LHI, xc);
  r = fma(-kd, DEXP_LLO, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double y = __ldg(&DEXP_TAB[k & 255]) * p;       // in [1, 2 + eps): adding k >> 8 (negative) to the exponent keeps it normal
  const double e = __hiloint2double(__double2hiint(y) + ((k >> 8) << 20), __double2loint(y));
  return x != x ? x : e;
}
__device__ __forceinline__ double tile_dlog(double s) {
  const int hi = __double2hiint(s), lo = __double2loint(s);
  const int i = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double e = (double)((hi >> 20) - 1023);
  const double inv = __ldg(&DLOG_TAB[2 * i]), lc = __ldg(&DLOG_TAB[2 * i + 1]);
  const double u = fma(m, inv, -1.0);
  double p = fma(u, 1.0 / 7.0, -1.0 / 6.0);
  p = fma(p, u, 0.2);
  p = fma(p, u, -0.25);
  p = fma(p, u, 1.0 / 3.0);
  p = fma(p, u, -0.5);
  const double r = fma(e, DLOG_LN2HI, lc) + fma(u * u, p, fma(e, DLOG_LN2LO, u));
  // anything but a positive normal number (integer tests on the high word, selects only): zero and subnormals (sums whose
  // every term underflowed) -> -inf, negative -> NaN, +inf and NaN -> themselves
  const bool normal = (unsigned)(hi - 0x00100000) < 0x7fe00000u;
  const double special = (unsigned)hi < 0x00100000u ? -CUDART_INF : hi < 0 ? CUDART_NAN : s + s;
  return normal ? r : special;
}
// exp of a non-positive argument / log of a positive sum in the sweep's precision
__device__ __forceinline__ float texp(float x) { return expf(x); }
__device__ __forceinline__ double texp(double x) { return tile_dexp_np(x); }
__device__ __forceinline__ float tlog(float x) { return logf(x); }
__device__ __forceinline__ double tlog(double x) { return tile_dlog(x); }

template <typename T> __device__ __forceinline__ void lds4(const T* p, T (&o)[4]);
template <> __device__ __forceinline__ void lds4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void lds4<double>(const double* p, double (&o)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T> __device__ __forceinline__ void sts4(T* p, const T (&v)[4]);
template <> __device__ __forceinline__ void sts4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void sts4<double>(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
template <typename T> __device__ __forceinline__ void ldg4(const T* p, T (&o)[4]);
template <> __device__ __forceinline__ void ldg4<float>(const float* p, float (&o)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ldg4<double>(const double* p, double (&o)[4]) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p + 2));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// L2-only loads (ld.global.cg): data another SM wrote during this kernel must not be served from a stale L1 line
template <typename T> __device__ __forceinline__ void ldcg4(const T* p, T (&o)[4]);
template <> __device__ __forceinline__ void ldcg4<float>(const float* p, float (&o)[4]) {
  const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ldcg4<double>(const double* p, double (&o)[4]) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p + 2));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// ---- TMA (cp.async.bulk.tensor) staging of a tile: one thread issues, an mbarrier (common.cuh) counts the bytes ----
__device__ __forceinline__ unsigned long long globaltimer_ns_tile() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// ---- per-cell backups on the four (already wall-resolved) neighbour values ----------------------------
// nu: logsumexp_a[log 1/4 + gamma*pa(s,a,nu) + rbar] with pa = w_off*S + dw*n_a and rbar = -step_cost away
// from the goal, i.e. kc + b1*S + a1*max + log sum_a exp(a1*(n_a - max)).
template <typename T, bool FAST>
__device__ __forceinline__ T nu_cell(const TileK<T>& k, T n0, T n1, T n2, T n3) {
  const T S = (n0 + n1) + (n2 + n3);
  // The largest exponent is exactly 0: sort out the maximum with a 5-comparator network and exponentiate only the
  // other three (one MUFU less; MUFU is the narrowest pipe of this kernel).
  const T h1 = fmax(n0, n1), l1 = fmin(n0, n1), h2 = fmax(n2, n3), l2 = fmin(n2, n3);
  const T m = fmax(h1, h2), mid = fmin(h1, h2);
  if constexpr (FAST) {
    const float t = -m * k.a2;
    const float sum = (1.0f + ex2_approx(fmaf(mid, k.a2, t))) + (ex2_approx(fmaf(l1, k.a2, t)) + ex2_approx(fmaf(l2, k.a2, t)));
    const float hi = fmaf(0.693147180559945f, lg2_approx(sum), fmaf(k.b1, S, fmaf(k.a1, m, k.kc)));
    return hi + fmaf(k.a1_lo, m, k.kc_lo);
  } else {
    const T sum = ((T)1 + texp(k.a1 * (mid - m))) + (texp(k.a1 * (l1 - m)) + texp(k.a1 * (l2 - m)));
    if (sizeof(T) == 4) return (k.kc + k.b1 * S + k.a1 * m + tlog(sum)) + (k.kc_lo + k.a1_lo * m);
    return k.kc + k.b1 * S + k.a1 * m + tlog(sum);
  }
}
// Two cells at once with the packed fp32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: two IEEE operations per lane and
// issue slot).  The kernel is bound by instruction issue, and 60 of the 141 instructions of a strip are adds, multiplies
// and FMAs; every packed lane performs exactly the scalar operation of nu_cell<float, true>, so the bits do not change.
#ifndef PIQL_TILE_PACKED
#define PIQL_TILE_PACKED 1
#endif
__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 f2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 fmax2(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
__device__ __forceinline__ float2 fmin2(float2 a, float2 b) { return make_float2(fminf(a.x, b.x), fminf(a.y, b.y)); }
__device__ __forceinline__ float2 ex2_2(float2 a) { return make_float2(ex2_approx(a.x), ex2_approx(a.y)); }
__device__ __forceinline__ float2 lg2_2(float2 a) { return make_float2(lg2_approx(a.x), lg2_approx(a.y)); }
__device__ __forceinline__ float2 nu_cell_fast2(const TileK<float>& k, float2 n0, float2 n1, float2 n2, float2 n3) {
  const float2 S = __fadd2_rn(__fadd2_rn(n0, n1), __fadd2_rn(n2, n3));
  const float2 h1 = fmax2(n0, n1), l1 = fmin2(n0, n1), h2 = fmax2(n2, n3), l2 = fmin2(n2, n3);
  const float2 m = fmax2(h1, h2), mid = fmin2(h1, h2);
  const float2 a2 = f2(k.a2);
  const float2 t = __fmul2_rn(m, f2(-k.a2));  // -m * a2
  const float2 e0 = ex2_2(__ffma2_rn(mid, a2, t)), e1 = ex2_2(__ffma2_rn(l1, a2, t)), e2 = ex2_2(__ffma2_rn(l2, a2, t));
  const float2 sum = __fadd2_rn(__fadd2_rn(f2(1.0f), e0), __fadd2_rn(e1, e2));
  const float2 hi = __ffma2_rn(f2(0.693147180559945f), lg2_2(sum), __ffma2_rn(f2(k.b1), S, __ffma2_rn(f2(k.a1), m, f2(k.kc))));
  return __fadd2_rn(hi, __ffma2_rn(f2(k.a1_lo), m, f2(k.kc_lo)));
}

constexpr int TILE_ERR_RANGE = 4;  // *err: a fast fp32 linear-domain sweep left the float range
// Tried and dropped (measured on B200, profiles/r2_vi_experiments.txt): nu in the exponent domain -- E = e^(a1 nu) kept beside nu
// in shared memory so that a sweep needs 2 MUFU per cell (lg2 of sum_a E(n_a), ex2 of the new value) instead of 4 and no
// maximum network.  Same instruction count (the 32 wall selects of a strip double), twice the shared memory per tile, and
// the exponentials lose the accuracy that arguments near zero give the maximum-shifted form: 3.74 against 3.23 us per sweep
// at 1026^2, 37.4 against 26.2 at 4098^2, and a uniform 1.4e-5 relative bias of the fixed point (the shifted form: 3e-6).

// eta = 1/2 damping of update_logz!: logaddexp(log(1/2) + new, log(1/2) + old)   value_iteration.jl:101-102
template <typename T, bool FAST> __device__ __forceinline__ T damp_half(T nw, T c) {
  const T d = fabs(nw - c), mx = fmax(nw, c);
  if constexpr (FAST) return fmaf(0.693147180559945f, lg2_approx(fmaf(0.5f, ex2_approx(-1.44269504088896f * d), 0.5f)), mx);
  else return mx + tlog((T)0.5 + (T)0.5 * texp(-d));
}
// logz: with Z = exp(logz), Z_new = sum_a exp(A_a) * (gamma * sum_a' P(a'|a) Z(n_a') + 1 - gamma), evaluated
// with the neighbour maximum m and M = max(m, 0) factored out (both exponents stay <= 0); ea_a = exp(A_a - amax).
template <typename T, bool FAST>
__device__ __forceinline__ T logz_cell(const TileK<T>& k, T c, T n0, T n1, T n2, T n3, T ea0, T ea1, T ea2, T ea3, T amax) {
  const T m = fmax(fmax(n0, n1), fmax(n2, n3));
  T e0, e1, e2, e3, em;
  if constexpr (FAST) {
    const float l2e = 1.44269504088896f, t = -m * l2e;
    e0 = ex2_approx(fmaf(n0, l2e, t)); e1 = ex2_approx(fmaf(n1, l2e, t));
    e2 = ex2_approx(fmaf(n2, l2e, t)); e3 = ex2_approx(fmaf(n3, l2e, t));
    em = ex2_approx(-fabsf(m) * l2e);
  } else {
    e0 = texp(n0 - m); e1 = texp(n1 - m); e2 = texp(n2 - m); e3 = texp(n3 - m);
    em = texp(-fabs(m));
  }
  const T wE = k.w_off * ((e0 + e1) + (e2 + e3));
  const bool pos = m >= (T)0;
  const T g1 = k.gamma * (pos ? (T)1 : em), g0 = k.omg * (pos ? em : (T)1);
  const T base = g1 * wE + g0, gd = g1 * k.dw;
  const T sum = (ea0 * (gd * e0 + base) + ea1 * (gd * e1 + base)) + (ea2 * (gd * e2 + base) + ea3 * (gd * e3 + base));
  if constexpr (FAST) {
    // new = M + amax + log(sum) and the eta = 1/2 damping log(1/2 e^new + 1/2 e^c) fused: c + log(1/2 + 1/2 sum e^(M+amax-c)),
    // one ex2 and one lg2 instead of lg2 + ex2 + lg2 (exponent clamped: 2^100 is far beyond any transient of the solve)
    const float r = sum * ex2_approx(fminf((fmaxf(m, 0.f) + amax - c) * 1.44269504088896f, 100.f));
    return fmaf(0.693147180559945f, lg2_approx(fmaf(0.5f, r, 0.5f)), c);
  } else {
    const T nw = fmax(m, (T)0) + amax + tlog(sum);
    return damp_half<T, FAST>(nw, c);
  }
}
// The fast fp32 form of the same recursion, relative to the cell's own value c instead of the neighbour maximum:
//   Z_new / e^c = sum_a e^(A_a) * (gamma * sum_a' P(a'|a) e^(n_a' - c) + (1 - gamma) e^(-c)),   stored = c + log(1/2 + 1/2 Z_new / e^c)
// (the eta = 1/2 damping fused as above).  Six MUFU per cell instead of seven -- the XU pipe bounds these sweeps -- and no sign
// cases; EA_a = e^(A_a) comes from the packed constant planes (tile_pack_aux_kernel folds max A in for this mode).  The sum is
// ~1 at the fixed point whatever the magnitude of c, so e^(-c) cannot overflow the products there; the exponents are clamped
// at 2^100 (four of them still add up far below FLT_MAX) so that a wild warm start stays finite and the damped iteration recovers.
// The clamp cannot bite at a fixed point: Z(s) >= (1 - gamma) sum_a e^(A_a), so -c <= -log(1 - gamma) - max_a A_a, a few tens at most.
#ifndef PIQL_LOGZ_OWN_SHIFT
#define PIQL_LOGZ_OWN_SHIFT 1
#endif
__device__ __forceinline__ float logz_cell_fast_own(const TileK<float>& k, float c, float n0, float n1, float n2, float n3,
                                                    float EA0, float EA1, float EA2, float EA3) {
  const float l2e = 1.44269504088896f, tc = -c * l2e;
  const float u0 = ex2_approx(fminf(fmaf(n0, l2e, tc), 100.f)), u1 = ex2_approx(fminf(fmaf(n1, l2e, tc), 100.f));
  const float u2 = ex2_approx(fminf(fmaf(n2, l2e, tc), 100.f)), u3 = ex2_approx(fminf(fmaf(n3, l2e, tc), 100.f));
  const float h = ex2_approx(fminf(tc, 100.f));
  const float base = fmaf(k.gamma * k.w_off, (u0 + u1) + (u2 + u3), k.omg * h), gd = k.gamma * k.dw;
  const float sum = (EA0 * fmaf(gd, u0, base) + EA1 * fmaf(gd, u1, base)) + (EA2 * fmaf(gd, u2, base) + EA3 * fmaf(gd, u3, base));
  return fmaf(0.693147180559945f, lg2_approx(fmaf(0.5f, sum, 0.5f)), c);
}
// ---- fast fp32 log Z in the linear domain ----------------------------------------------------------------------------
// With Z = e^(logZ) the recursion is linear, Z_new = sum_a e^(A_a) (gamma sum_a' P(a'|a) Z(n_a') + 1 - gamma), and so is the
// reference's eta = 1/2 damping: log(1/2 e^new + 1/2 e^old) is the log of 1/2 Z_new + 1/2 Z_old (value_iteration.jl:101-102).
// Exchanging the two sums, sum_a e^(A_a) sum_a' P(a'|a) Z(n_a') = sum_a' M_a' Z(n_a') with M_a' = w_off sum_a e^(A_a) + dw e^(A_a'),
//     Z_stored = 1/2 Z + sum_a' G_a' Z(n_a') + C,      G_a' = 1/2 gamma M_a',   C = 1/2 (1 - gamma) sum_a e^(A_a):
// five FMAs per cell on four constant planes (C follows from them), NO transcendental at all in a sweep (the log-domain forms need 6-7 MUFU per
// cell on the pipe that bounds them) and no scalar constants whose fp32 rounding would act as a 1/(1-gamma)-amplified
// bias: the planes are computed in double and rounded once per cell.  All terms are positive, so nothing cancels.  The
// padded iterate holds Z for the whole solve (packing exponentiates, unpacking takes the log, the stop rule compares logs);
// the values do not depend on the tiling or the GPU count.  Range: e^(logZ) must stay a normal float, |logZ| < ~87, checked
// at the write-back (TILE_ERR_RANGE; PIQL_F32 / PIQL_F64 keep the log-domain form).
#ifndef PIQL_LOGZ_ZDOM
#define PIQL_LOGZ_ZDOM 1
#endif
template <typename T, int KIND, bool FAST> constexpr bool kZDom = PIQL_LOGZ_ZDOM && KIND == KIND_LOGZ && FAST && sizeof(T) == 4;
// The two linear fast fp32 sweeps (log Z in the linear domain, gap) are FMA chains on per-cell constant planes; read from
// global memory in every sub-sweep the planes make them L2-bandwidth bound (4-5 planes x 4 B per cell and sweep, ~10 TB/s
// L2 -> SM at 4098^2).  So the planes of a tile are staged in shared memory beside the iterate, once per round (TMA, one 3-D
// box per plane; once per launch in the persistent kernel): global traffic per cell and round falls from 4 x 16..20 B to
// 16..20 B.  log Z needs only its four G planes: sum_a' G_a' = 1/2 gamma sum_a e^(A_a), so C = (1 - gamma) / gamma * sum_a' G_a'.
#ifndef PIQL_TILE_PLANES
#define PIQL_TILE_PLANES 1
#endif
template <typename T, int KIND, bool FAST, bool PL>  // PL: the plan stages the planes (TilePlan::planes > 0)
constexpr int kPlanes = !(PL && PIQL_TILE_PLANES && FAST && sizeof(T) == 4) ? 0 : kZDom<T, KIND, FAST> ? 4 : KIND == KIND_GAP ? 5 : 0;
__device__ __forceinline__ float2 logz_cell_lin2(float2 c, float2 z0, float2 z1, float2 z2, float2 z3, float2 G0, float2 G1,
                                                 float2 G2, float2 G3, float kappa) {
  // C = kappa * (G0 + G1 + G2 + G3), kappa = (1 - gamma) / gamma (C is ~1 % of the result: kappa's rounding is harmless)
  const float2 C = __fmul2_rn(f2(kappa), __fadd2_rn(__fadd2_rn(G0, G1), __fadd2_rn(G2, G3)));
  return __ffma2_rn(G0, z0, __ffma2_rn(G1, z1, __ffma2_rn(G2, z2, __ffma2_rn(G3, z3, __ffma2_rn(f2(0.5f), c, C)))));
}

// gap: D + sum_a pi_a * pa(s,a,gap), no discount on the gap term   value_iteration.jl:165
template <typename T>
__device__ __forceinline__ T gap_cell(const TileK<T>& k, T n0, T n1, T n2, T n3, T p0, T p1, T p2, T p3, T D) {
  const T wS = k.w_off * ((n0 + n1) + (n2 + n3));
  return D + ((p0 * (wS + k.dw * n0) + p1 * (wS + k.dw * n1)) + (p2 * (wS + k.dw * n2) + p3 * (wS + k.dw * n3)));
}

// |new - old| of the reference's stop rule from the two last iterates.  logz stores the damped value:
// stored = log(1/2 e^new + 1/2 e^old)  =>  new - old = log(2 e^(stored-old) - 1).
// ... and in the linear domain: Z_stored = 1/2 Z_new + 1/2 Z_old  =>  new - old = log((2 Z_stored - Z_old) / Z_old).
__device__ __forceinline__ double sweep_diff_z(double stored, double old) {
  const double r = (2.0 * stored - old) / old;
  return r > 0.0 ? fabs(log(r)) : CUDART_INF;
}
template <int KIND> __device__ __forceinline__ double sweep_diff(double stored, double old) {
  const double d = stored - old;
  if (KIND != KIND_LOGZ) return fabs(d);
  const double x = 2.0 * expm1(d);
  return x > -1.0 ? fabs(log1p(x)) : CUDART_INF;
}

// The per-cell work of one 128-cell strip, on registers: lane l holds rows 4l .. 4l+3 of the column (v), of its left and right
// neighbours (L, R), the cells above row 4l and below row 4l+3 (up, dn), the four cell bytes (mk) and the strip's constants
// (ax: planes of the linear sweeps / A_a / pi_a, D).  Every cell by the generic formula (the goal and its neighbours are
// redone by the callers).  Shared by every schedule of the sweeps -- one launch per round, persistent tiles, resident
// columns -- so that their values agree bit for bit.
template <typename T, int KIND, bool FAST>
__device__ __forceinline__ void strip_compute(const TileK<T>& k, const T (&v)[4], const T (&L)[4], const T (&R)[4], T up, T dn,
                                              uint32_t mk, const T (&ax)[5][4], T (&out)[4]) {
  if constexpr (kZDom<T, KIND, FAST>) {
    float nb[4][4];  // [action][row]
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t b = mk >> (8 * r);
      const float c = v[r];
      nb[0][r] = (b & 1u) ? c : (r < 3 ? v[r < 3 ? r + 1 : 3] : dn);
      nb[1][r] = (b & 2u) ? c : (r > 0 ? v[r > 0 ? r - 1 : 0] : up);
      nb[2][r] = (b & 4u) ? c : R[r];
      nb[3][r] = (b & 8u) ? c : L[r];
    }
#pragma unroll
    for (int r = 0; r < 4; r += 2) {
      const float2 o = logz_cell_lin2(f2(v[r], v[r + 1]), f2(nb[0][r], nb[0][r + 1]), f2(nb[1][r], nb[1][r + 1]),
                                      f2(nb[2][r], nb[2][r + 1]), f2(nb[3][r], nb[3][r + 1]), f2(ax[0][r], ax[0][r + 1]),
                                      f2(ax[1][r], ax[1][r + 1]), f2(ax[2][r], ax[2][r + 1]), f2(ax[3][r], ax[3][r + 1]), k.omg);
      out[r] = o.x; out[r + 1] = o.y;
    }
  } else if constexpr (PIQL_TILE_PACKED && KIND == KIND_NU && FAST && sizeof(T) == 4) {
    // rows (0,1) and (2,3) as packed pairs
    float nb[4][4];  // [action][row]
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t b = mk >> (8 * r);
      const float c = v[r];
      nb[0][r] = (b & 1u) ? c : (r < 3 ? v[r < 3 ? r + 1 : 3] : dn);
      nb[1][r] = (b & 2u) ? c : (r > 0 ? v[r > 0 ? r - 1 : 0] : up);
      nb[2][r] = (b & 4u) ? c : R[r];
      nb[3][r] = (b & 8u) ? c : L[r];
    }
#pragma unroll
    for (int r = 0; r < 4; r += 2) {
      const float2 o = nu_cell_fast2(k, f2(nb[0][r], nb[0][r + 1]), f2(nb[1][r], nb[1][r + 1]), f2(nb[2][r], nb[2][r + 1]),
                                     f2(nb[3][r], nb[3][r + 1]));
      out[r] = o.x; out[r + 1] = o.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t b = mk >> (8 * r);
      const T c = v[r];
      // f(take_step(s,a)): the neighbour, or the cell itself when a wall blocks the move  GridWorld.jl:29-32
      const T n0 = (b & 1u) ? c : (r < 3 ? v[r < 3 ? r + 1 : 3] : dn);
      const T n1 = (b & 2u) ? c : (r > 0 ? v[r > 0 ? r - 1 : 0] : up);
      const T n2 = (b & 4u) ? c : R[r];
      const T n3 = (b & 8u) ? c : L[r];
      T o;
      if (KIND == KIND_NU) o = nu_cell<T, FAST>(k, n0, n1, n2, n3);
      else if constexpr (PIQL_LOGZ_OWN_SHIFT && KIND == KIND_LOGZ && FAST) o = logz_cell_fast_own(k, c, n0, n1, n2, n3, ax[0][r], ax[1][r], ax[2][r], ax[3][r]);
      else if (KIND == KIND_LOGZ) o = logz_cell<T, FAST>(k, c, n0, n1, n2, n3, ax[0][r], ax[1][r], ax[2][r], ax[3][r], ax[4][r]);
      else o = gap_cell<T>(k, n0, n1, n2, n3, ax[0][r], ax[1][r], ax[2][r], ax[3][r], ax[4][r]);
      out[r] = o;
    }
  }
}
// planes a sweep reads per cell when they come from global memory (0 for nu)
template <typename T, int KIND, bool FAST>
constexpr int kPlanesGlobal = KIND == KIND_NU ? 0 : (kZDom<T, KIND, FAST> || (PIQL_LOGZ_OWN_SHIFT && KIND == KIND_LOGZ && FAST)) ? 4 : 5;

// One column of one Jacobi sweep of the staged tile: lane l recomputes rows 4l .. 4l+3 of column jj from src into dst,
// every cell by the generic formula (the goal and its neighbours are redone by tile_goal_fixup).
template <typename T, int KIND, bool FAST, bool PL>
__device__ __forceinline__ void tile_strip(const TileK<T>& k, const T* __restrict__ src, T* __restrict__ dst,
                                           const uint32_t* __restrict__ msk, const T* __restrict__ aux, int org, int lane, int jj) {
  T v[4], L[4], R[4], out[4];
  lds4(src + jj * TEI + 4 * lane, v);
  lds4(src + (jj - 1) * TEI + 4 * lane, L);
  lds4(src + (jj + 1) * TEI + 4 * lane, R);
  const uint32_t mk = msk[jj * TMP + lane];
  // rows 0 and 127 have no vertical neighbour in the tile: they are halo, any value will do
  const T up = __shfl_up_sync(FULL, v[3], 1), dn = __shfl_down_sync(FULL, v[0], 1);
  T ax[5][4];
  if constexpr (kPlanes<T, KIND, FAST, PL> > 0) {
    // the tile's planes sit behind the ping-pong pair in shared memory (aux points at them)
#pragma unroll
    for (int pl = 0; pl < kPlanes<T, KIND, FAST, PL>; pl++) lds4(aux + (pl * k.EJ + jj) * TEI + 4 * lane, ax[pl]);
  } else if (KIND != KIND_NU) {
    const int sidx = org + jj * k.ld + 4 * lane;
#pragma unroll
    for (int pl = 0; pl < kPlanesGlobal<T, KIND, FAST>; pl++) ldg4(aux + (size_t)pl * k.np + sidx, ax[pl]);
  }
  strip_compute<T, KIND, FAST>(k, v, L, R, up, dn, mk, ax, out);
  sts4(dst + jj * TEI + 4 * lane, out);
}

// Jacobi sweep number `sub` (0-based) of the staged tile, src -> dst, all 128 rows; the caller puts a CTA barrier after it.
// src holds sweep `sub` of the round in columns sub .. EJ-1-sub only (the valid region loses one ring per sweep), so the
// columns worth computing are 1+sub .. EJ-2-sub: the trapezoid, not the rectangle.  What lies outside keeps older values
// that no kept cell ever reads.
template <typename T, int KIND, bool FAST, bool PL>
__device__ __forceinline__ void tile_subsweep(const TileK<T>& k, const T* __restrict__ src, T* __restrict__ dst,
                                              const uint32_t* __restrict__ msk, const T* __restrict__ aux, int org, int lane,
                                              int warp, int NW, int sub) {
#pragma unroll TILE_UNROLL
  for (int jj = 1 + sub + warp; jj < k.EJ - 1 - sub; jj += NW) tile_strip<T, KIND, FAST, PL>(k, src, dst, msk, aux, org, lane, jj);
}

// The goal (terminal value) and, for nu, its four neighbours (a step into the goal earns the reward, so the
// average reward is not the constant the generic formula assumes) are recomputed from src by five threads.
template <typename T, int KIND, bool FAST>
__device__ __noinline__ void tile_goal_fixup(const TileK<T>& k, const T* __restrict__ src, T* __restrict__ dst,
                                             const uint32_t* __restrict__ msk, int gi, int gj) {
  const int t = threadIdx.x;  // 0: goal, 1..4: (+1,0) (-1,0) (0,+1) (0,-1)
  if (t >= (KIND == KIND_NU ? 5 : 1)) return;
  const int i = gi + (t == 1) - (t == 2), j = gj + (t == 3) - (t == 4);
  if (i < 0 || i >= TEI || j < 1 || j >= k.EJ - 1) return;  // not a cell this tile computes
  const uint32_t b = (msk[j * TMP + (i >> 2)] >> (8 * (i & 3))) & 0xffu;
  const T c = src[j * TEI + i];
  T o;
  if (t == 0) {
    // terminal_condition: nu, gap -> 0; logz -> 0 - V(goal), damped like every other state
    if constexpr (kZDom<T, KIND, FAST>) o = (T)0.5 * k.term + (T)0.5 * c;  // linear domain: term = e^(-V(goal))
    else o = (KIND == KIND_LOGZ) ? damp_half<T, FAST>(k.term, c) : k.term;
  } else {
    if (b & CELL_WALL) return;
    // rows 0 / 127 have no vertical neighbour in the tile; they are halo (never kept), any value will do
    T nb[NA];
    nb[0] = (b & 1u) || i + 1 >= TEI ? c : src[j * TEI + i + 1];
    nb[1] = (b & 2u) || i - 1 < 0 ? c : src[j * TEI + i - 1];
    nb[2] = (b & 4u) ? c : src[(j + 1) * TEI + i];
    nb[3] = (b & 8u) ? c : src[(j - 1) * TEI + i];
    const int cidx = k.goal_p + (t == 1) - (t == 2) + k.ld * ((t == 3) - (t == 4));
    o = backup_nu(k.k, cidx, b, nb);
  }
  dst[j * TEI + i] = o;
}

// linear-domain log Z: Z at a free cell must be a normal positive float (else the sweep left its range)
template <typename T> __device__ __forceinline__ bool strip_not_normal(const T (&f)[4], uint32_t mk) {
  bool bad = false;
#pragma unroll
  for (int r = 0; r < 4; r++) bad |= !((mk >> (8 * r)) & CELL_WALL) && !((double)f[r] >= 1.1754944e-38 && (double)f[r] <= 3.0e38);
  return bad;
}

// ---- the kernel ------------------------------------------------------------------------------------------
// grid = tiles_i * tiles_j CTAs of NW warps; dynamic shared memory = EJ * (2 * 128 * sizeof(T) + 144) bytes.
template <typename T, int KIND, bool FAST, bool PL>
__global__ void __launch_bounds__(sizeof(T) == 4 ? PIQL_TILE_F32_MAXT : PIQL_TILE_F64_MAXT, sizeof(T) == 4 ? PIQL_TILE_F32_MINB : PIQL_TILE_F64_MINB)
tile_sweep_kernel(const TileK<T> k, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmM,
                  const __grid_constant__ CUtensorMap tmP, int use_tma, const uint8_t* __restrict__ maskp, const T* __restrict__ X,
                  T* __restrict__ Y, const T* __restrict__ auxg, int nsub, int want_res, double* __restrict__ partial,
                  const TileHalo halo) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ double red[32];
  __shared__ __align__(8) unsigned long long mbar;
  const int EJ = k.EJ, ld = k.ld;
  constexpr int NP = kPlanes<T, KIND, FAST, PL>;
  T* bufA = reinterpret_cast<T*>(tile_smem);
  T* bufB = bufA + EJ * TEI;
  T* planes = bufB + EJ * TEI;  // NP constant planes of the tile (fast fp32 log Z / gap)
  uint32_t* msk0 = reinterpret_cast<uint32_t*>(planes + NP * EJ * TEI);
  const T* aux = NP > 0 ? planes : auxg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
  // column slabs put their two edge tile columns first in launch order: their push is what the neighbours wait for
  const int ti = blockIdx.x % k.tiles_i, tc = blockIdx.x / k.tiles_i;
  const int tj = halo.tj_last < 0 ? tc + k.tj0 : tc == 0 ? halo.tj_first : tc == 1 ? halo.tj_last : halo.tj_first + tc - 1;
  const int org = ti * k.tti + ld * (tj * k.TJ);  // padded index of the tile's (row 0, column 0)
  const int moff = (ti * k.tti) & 15;              // bytes the staged cell-byte box starts before row 0 (0 or 8)
  const uint32_t* msk = msk0 + (moff >> 2);

  // Programmatic dependent launch: this grid may be scheduled while the previous launch (which wrote X) drains;
  // let the next one do the same, then wait for the previous grid's memory before the first global read.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // multi-GPU column slabs: the tiles of the slab's edge columns read halo columns the neighbouring rank pushes
  const bool edge_l = halo.out_left != nullptr && tj == halo.tj_first, edge_r = halo.out_right != nullptr && tj == halo.tj_last;
  if (halo.wait_epoch && (edge_l || edge_r)) {
    if (threadIdx.x == 0) {
      const unsigned long long t0 = globaltimer_ns_tile();
      for (int side = 0; side < 2; side++) {
        const volatile unsigned* f = side ? halo.in_right : halo.in_left;
        if (!f || !(side ? edge_r : edge_l)) continue;
        while (*f < halo.wait_epoch) {
          if (globaltimer_ns_tile() - t0 > 10000000000ull) { *halo.err = 2; break; }
          __nanosleep(100);
        }
      }
      if (halo.wait_stats) {
        atomicAdd(halo.wait_stats, globaltimer_ns_tile() - t0);
        atomicAdd(halo.wait_stats + 1, 1ull);
      }
      __threadfence_system();
      asm volatile("fence.proxy.async;" ::: "memory");  // the peer's stores, just acquired, before this thread's TMA reads
    }
    if (!use_tma) __syncthreads();
  }

  // stage the tile (every access in bounds thanks to the ghost frame)
  if (use_tma) {
    // two TMA boxes of 128 x EJ elements (iterate, cell bytes) land in shared memory; nobody spends an issue slot
    // or a register on the copy, and the whole tile is in flight at once
    if (threadIdx.x == 0) {
      mbar_init(&mbar, 1);
      mbar_expect_tx(&mbar, (unsigned)(EJ * ((1 + NP) * TEI * sizeof(T) + TMB)));
      tma_load_2d(bufA, &tmX, ti * k.tti, tj * k.TJ, &mbar);
      tma_load_2d(msk0, &tmM, ti * k.tti - moff, tj * k.TJ, &mbar);  // rows past the array's end are zero-filled
#pragma unroll
      for (int q = 0; q < NP; q++) tma_load_3d(planes + q * EJ * TEI, &tmP, ti * k.tti, tj * k.TJ, q, &mbar);
    }
    __syncthreads();  // the barrier word is initialised before anyone polls it
    mbar_wait(&mbar, 0);
    // the outermost columns are never recomputed: keep both ping-pong copies defined
    if (warp < 2) {
      const int jj = warp == 0 ? 0 : EJ - 1;
      T v[4];
      lds4(bufA + jj * TEI + 4 * lane, v);
      sts4(bufB + jj * TEI + 4 * lane, v);
    }
  } else {
    for (int jj = warp; jj < EJ; jj += NW) {
      const int g = org + jj * ld + 4 * lane;
      T v[4];
      ldg4(X + g, v);
      sts4(bufA + jj * TEI + 4 * lane, v);
      if (jj == 0 || jj == EJ - 1) sts4(bufB + jj * TEI + 4 * lane, v);
      msk0[jj * TMP + (moff >> 2) + lane] = __ldg(reinterpret_cast<const uint32_t*>(maskp + g));
#pragma unroll
      for (int q = 0; q < NP; q++) {
        ldg4(auxg + (size_t)q * k.np + g, v);
        sts4(planes + (q * EJ + jj) * TEI + 4 * lane, v);
      }
    }
  }
  __syncthreads();

  // is the goal (or one of its neighbours) inside this tile?  block-uniform
  const int gi = k.goal_pi - ti * k.tti, gj = k.goal_pj - tj * k.TJ;
  const bool near_goal = gi >= -1 && gi <= TEI && gj >= -1 && gj <= EJ;

  T* src = bufA;
  T* dst = bufB;
  for (int sub = 0; sub < nsub; sub++) {
    tile_subsweep<T, KIND, FAST, PL>(k, src, dst, msk, aux, org, lane, warp, NW, sub);
    __syncthreads();
    if (near_goal) {  // block-uniform: the tile(s) around the goal redo those cells after every sweep
      tile_goal_fixup<T, KIND, FAST>(k, src, dst, msk, gi, gj);
      __syncthreads();
    }
    T* t = src; src = dst; dst = t;
  }

  // write the interior back (src: last iterate, dst: the one before); residual of the last sweep
  double res = 0.0;
  bool bad = false;
  if (lane >= k.kh / 4 && lane < 32 - k.kh / 4) {
    for (int jj = k.kh + warp; jj < k.kh + k.TJ; jj += NW) {
      T f[4];
      lds4(src + jj * TEI + 4 * lane, f);
      if constexpr (kZDom<T, KIND, FAST>) bad |= strip_not_normal(f, msk[jj * TMP + lane]);
      const int g = org + jj * ld + 4 * lane;
      if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(Y + g) = make_float4((float)f[0], (float)f[1], (float)f[2], (float)f[3]);
      } else {
        *reinterpret_cast<double2*>(Y + g) = make_double2((double)f[0], (double)f[1]);
        *reinterpret_cast<double2*>(Y + g + 2) = make_double2((double)f[2], (double)f[3]);
      }
      // the slab's first / last TILE_KH interior columns are the neighbouring rank's halo: peer stores over NVLink
      T* peer = (edge_l && jj < 2 * k.kh) ? static_cast<T*>(halo.out_left)
                                         : (edge_r && jj >= k.TJ) ? static_cast<T*>(halo.out_right) : nullptr;
      if (peer) sts4(peer + g, f);
      if (want_res) {
        T p[4];
        lds4(dst + jj * TEI + 4 * lane, p);
        const uint32_t mk = msk[jj * TMP + lane];
#pragma unroll
        for (int r = 0; r < 4; r++)
          if (!((mk >> (8 * r)) & CELL_WALL))
            res += kZDom<T, KIND, FAST> ? sweep_diff_z((double)f[r], (double)p[r]) : sweep_diff<KIND>((double)f[r], (double)p[r]);
      }
    }
  }
  if (kZDom<T, KIND, FAST> && bad) *halo.err = TILE_ERR_RANGE;
  if (want_res) {  // uniform
    const double s = block_sum(res, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
  if (edge_l || edge_r) {  // block-uniform: the last edge CTA of a side to finish raises that neighbour's counter
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int side = 0; side < 2; side++) {
        if (!(side ? edge_r : edge_l)) continue;
        if (atomicAdd(halo.done + side, 1u) == (unsigned)k.tiles_i - 1u) {
          halo.done[side] = 0u;
          __threadfence_system();
          *(volatile unsigned*)(side ? halo.flag_right : halo.flag_left) = halo.post_epoch;
        }
      }
    }
  }
}

// ---- the persistent variant: all rounds of a batch of sweeps in ONE cooperative launch ---------------------------
// For grids whose tiles are all resident at once (<= CTAs per SM x SMs: the 10^6-cell problems), a launch per
// TILE_KH sweeps pays a grid-wide barrier, a tail and a pipeline refill every few microseconds.  Here every CTA keeps
// its tile for the whole batch and synchronises with its (up to 8) neighbouring tiles only: round r waits until they
// have published round r-1 (a counter per tile, release/acquire at gpu scope).  The tile's interior STAYS in shared memory
// from round to round: only the halo ring (kh cells deep) is re-read from the ping-pong global arrays, and only the ring of
// interior cells the neighbours need (kh deep) is written back -- the whole interior just once, after the last round.  (The
// first version re-staged the whole tile with TMA and wrote the whole interior back every round: it never beat one launch
// per round.)  Waiting for "neighbours finished r-1" before round r also makes the write-back of round r safe: it
// overwrites the copy they read in round r-1.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T, int KIND, bool FAST, bool PL>
__global__ void __launch_bounds__(sizeof(T) == 4 ? PIQL_TILE_F32_MAXT : PIQL_TILE_F64_MAXT, sizeof(T) == 4 ? PIQL_TILE_F32_MINB : PIQL_TILE_F64_MINB)
tile_persist_kernel(const TileK<T> k, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmP, T* __restrict__ XA,
                    T* __restrict__ XB, const T* __restrict__ auxg,
                    int nrounds, int nsub_last, int want_res, double* __restrict__ partial, unsigned* __restrict__ flags,
                    int tiles_j, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ double red[32];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int s_abort;
  const int EJ = k.EJ, ld = k.ld;
  constexpr int NP = kPlanes<T, KIND, FAST, PL>;
  T* bufA = reinterpret_cast<T*>(tile_smem);
  T* bufB = bufA + EJ * TEI;
  T* planes = bufB + EJ * TEI;
  uint32_t* msk0 = reinterpret_cast<uint32_t*>(planes + NP * EJ * TEI);
  const T* aux = NP > 0 ? planes : auxg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
  const int ti = blockIdx.x % k.tiles_i, tj = blockIdx.x / k.tiles_i;
  const int org = ti * k.tti + ld * (tj * k.TJ);
  const int moff = (ti * k.tti) & 15;
  const uint32_t* msk = msk0 + (moff >> 2);
  const int gi = k.goal_pi - ti * k.tti, gj = k.goal_pj - tj * k.TJ;
  const bool near_goal = gi >= -1 && gi <= TEI && gj >= -1 && gj <= EJ;

  if (threadIdx.x == 0) mbar_init(&mbar, 1);
  __syncthreads();
  T* src = bufA;
  T* dst = bufB;
  for (int r = 0; r < nrounds; r++) {
    const int nsub = r == nrounds - 1 ? nsub_last : k.kh;
    const T* Xin = (r & 1) ? XB : XA;
    T* Xout = (r & 1) ? XA : XB;
    // neighbours (8-connected: a 4-deep halo reaches into the diagonal tiles) have finished round r-1
    if (r > 0 && threadIdx.x < 8) {
      const int q = threadIdx.x < 4 ? threadIdx.x : threadIdx.x + 1;  // skip the centre of the 3 x 3 stencil
      const int ni = ti + q % 3 - 1, nj = tj + q / 3 - 1;
      if (ni >= 0 && ni < k.tiles_i && nj >= 0 && nj < tiles_j) {
        const unsigned* f = flags + ni + k.tiles_i * nj;
        const unsigned long long t0 = globaltimer_ns_tile();
        while (ld_acquire_gpu(f) < (unsigned)r) {
          if (*(volatile int*)err) break;  // somebody else gave up (or left the linear-domain range): so do we
          if (globaltimer_ns_tile() - t0 > 5000000000ull) { *err = 1; break; }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_abort = *(volatile int*)err;
    __syncthreads();
    if (s_abort) break;  // block-uniform: somebody timed out; the host reports it
    const int hl = k.kh / 4;  // halo lanes at the top / bottom of a column (a lane holds 4 rows)
    if (r == 0) {
      // the whole tile, its cell bytes and the constant planes: TMA boxes, once per launch
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_expect_tx(&mbar, (unsigned)(EJ * (TEI * sizeof(T) + TMB + NP * TEI * sizeof(T))));
        tma_load_2d(bufA, &tmA, ti * k.tti, tj * k.TJ, &mbar);
        tma_load_2d(msk0, &tmM, ti * k.tti - moff, tj * k.TJ, &mbar);
#pragma unroll
        for (int q = 0; q < NP; q++) tma_load_3d(planes + q * EJ * TEI, &tmP, ti * k.tti, tj * k.TJ, q, &mbar);
      }
      mbar_wait(&mbar, 0);
      if (warp < 2) {  // the outermost columns are never recomputed: keep both ping-pong copies defined
        const int jj = warp == 0 ? 0 : EJ - 1;
        T v[4];
        lds4(bufA + jj * TEI + 4 * lane, v);
        sts4(bufB + jj * TEI + 4 * lane, v);
      }
    } else {
      // the interior is already here (bufA: every round but the last has an even number of sweeps); refresh the halo ring
      // from what the neighbours published: whole columns at the left / right edge, the top / bottom strips elsewhere
      for (int jj = warp; jj < EJ; jj += NW) {
        const bool edge = jj < k.kh || jj >= EJ - k.kh;
        if (edge || lane < hl || lane >= 32 - hl) {
          T v[4];
          ldcg4(Xin + org + jj * ld + 4 * lane, v);
          sts4(bufA + jj * TEI + 4 * lane, v);
          if (jj == 0 || jj == EJ - 1) sts4(bufB + jj * TEI + 4 * lane, v);
        }
      }
    }
    __syncthreads();
    src = bufA; dst = bufB;
    for (int sub = 0; sub < nsub; sub++) {
      tile_subsweep<T, KIND, FAST, PL>(k, src, dst, msk, aux, org, lane, warp, NW, sub);
      __syncthreads();
      if (near_goal) {
        tile_goal_fixup<T, KIND, FAST>(k, src, dst, msk, gi, gj);
        __syncthreads();
      }
      T* t = src; src = dst; dst = t;
    }
    // publish: after the last round the whole interior, before that only the ring of interior cells within kh of its
    // boundary -- what the eight neighbours read as their halo
    const bool last_round = r == nrounds - 1;
    if (lane >= hl && lane < 32 - hl) {
      for (int jj = k.kh + warp; jj < k.kh + k.TJ; jj += NW) {
        const bool ring = jj < 2 * k.kh || jj >= k.TJ || lane < 2 * hl || lane >= 32 - 2 * hl;
        if (!last_round && !ring) continue;
        T f[4];
        lds4(src + jj * TEI + 4 * lane, f);
        if constexpr (kZDom<T, KIND, FAST>) { if (strip_not_normal(f, msk[jj * TMP + lane])) *err = TILE_ERR_RANGE; }
        const int g = org + jj * ld + 4 * lane;
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(Xout + g) = make_float4((float)f[0], (float)f[1], (float)f[2], (float)f[3]);
        } else {
          *reinterpret_cast<double2*>(Xout + g) = make_double2((double)f[0], (double)f[1]);
          *reinterpret_cast<double2*>(Xout + g + 2) = make_double2((double)f[2], (double)f[3]);
        }
      }
    }
    __threadfence();
    __syncthreads();  // every thread's stores are fenced; also: all reads of src/dst done before the next TMA overwrite
    if (threadIdx.x == 0) st_release_gpu(flags + blockIdx.x, (unsigned)(r + 1));
  }
  // residual of the last sweep (src: last iterate, dst: the one before)
  if (want_res) {
    double res = 0.0;
    if (lane >= k.kh / 4 && lane < 32 - k.kh / 4) {
      for (int jj = k.kh + warp; jj < k.kh + k.TJ; jj += NW) {
        T f[4], p[4];
        lds4(src + jj * TEI + 4 * lane, f);
        lds4(dst + jj * TEI + 4 * lane, p);
        const uint32_t mk = msk[jj * TMP + lane];
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (!((mk >> (8 * q)) & CELL_WALL))
            res += kZDom<T, KIND, FAST> ? sweep_diff_z((double)f[q], (double)p[q]) : sweep_diff<KIND>((double)f[q], (double)p[q]);
      }
    }
    const double s = block_sum(res, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

// ---- the column-resident variant: every cell computed exactly once per sweep ---------------------------------------------
// Temporal blocking pays for its few launches with halo recomputation: at 1026^2 (8 sweeps per round, 128 x 40 tiles) a CTA
// computes 1.9 cells per cell it keeps.  When a sweep is at most two 128-row strips per warp (28 warps x 148 SMs: ~10^6 cells, the
// config-4 grid) there is a schedule with no redundancy at all: one CTA per SM owns a slab of whole columns (all rows of the
// free cells' bounding box: 8 strips of 128 rows at 1026^2, so no vertical halo either -- the border is wall, grid_load checks
// it), keeps it in shared memory for the whole batch of sweeps, and exchanges ONE column with each neighbour per sweep through an
// L2 mailbox:
//   * every 32-bit word travels as a 64-bit {word, sequence number} pair written and polled with single 64-bit accesses
//     (single-copy atomic), so there is no fence, no flag round trip and no grid barrier: a reader spins on the data itself
//     (one way: 0.3 us between two otherwise idle SMs, 0.65 us with every SM pair talking; tools/probe/mail_latency.cu);
//   * a warp owns its (at most two) strips for the whole launch, so what does not change lives in its registers: the strip's own
//     values (stored for the neighbours, never reloaded), the cell bytes and, for the linear fast sweeps, the constant planes --
//     a strip of those is four shared-memory accesses instead of nine, and the shared-memory pipe is what bounds them;
//   * a warp's first strip is an edge-column strip while they last: it is posted the moment it is stored; the neighbours' strips
//     of the same sweep are collected at its end (into the halo columns of the buffer the next sweep reads) by the warps the work
//     list names: those whose strips are all interior first -- they finish ~1000 cycles before the owners of edge strips, and a
//     fetch is another ~1000 -- two strips' loads in flight at once;
//   * mailboxes are double-buffered by sweep parity (a neighbour can be at most one sweep ahead);
//   * the goal cell and (nu) its neighbours are patched, one lane per cell, by the warp that owns them before anything is posted.
// Per sweep and CTA: one __syncthreads.  The per-cell arithmetic is strip_compute, shared with the other schedules: same bits.
// What bounds it (clock64 stamps, PIQL_TILE_COLRES_TRACE; nu at 1026^2): two strips per warp take 3400-3900 cycles, a fetch of
// the neighbours' strips ~1000 (a global round trip under this kernel's shared-memory load), barrier and loop head ~600: 2.7 us
// per sweep against 3.05 for the tiled kernel (nu), 2.2 against 2.8 (log Z), 2.65 against 2.95 (gap); 2.3 / 1.4 / 1.8 without
// the exchange.  DESIGN.md 5.4 lists what was tried against that latency.
struct ColGeom {
  int r0;           // first padded row of the resident columns (a multiple of 4: vector accesses stay aligned)
  int NS, RS;       // 128-row strips per column, RS = 128 * NS rows
  int c0;           // first padded column of CTA 0
  int wbase, wrem;  // CTA c owns wbase + (c < wrem) columns
  int wmax;         // widest slab (shared-memory layout)
  int ncta;
  int nwc;          // warps per CTA
  int G;            // strips per warp in the work list (kColSlots: 2 in fp32, 4 in Float64)
  int diag;         // timing experiments only (PIQL_TILE_COLRES_DIAG): 1 = no exchange at all (wrong values at the slab seams)
  int diag_cta;     // the CTA PIQL_TILE_COLRES_TRACE follows
  int diag_sw;      // first of the four sweeps it stamps (PIQL_TILE_COLRES_TRACE_SW, default 10)
};
constexpr int COL_ERR_TIMEOUT = 1;
// 28 warps of 72 registers in fp32 (the planes of the linear sweeps live in registers), 16 of 128 in Float64
template <typename T> constexpr int kColMaxThreads = sizeof(T) == 8 ? 512 : 896;
constexpr uint32_t COL_NONE = 0xffffffffu;  // host work list: no item
constexpr uint32_t COL_END = 0u;           // shared-memory work list: no item (offset 0 is the halo column, never a strip)
// work-list item in shared memory: offset of the strip's first row in a buffer (20 bits) and flags
constexpr uint32_t COL_OFS = 0xfffffu, COL_EDGE_L = 1u << 20, COL_EDGE_R = 1u << 21, COL_HAS_UP = 1u << 22, COL_HAS_DN = 1u << 23,
                   COL_GOAL = 1u << 24;
__device__ __forceinline__ void ld_mail(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_mail(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
template <typename T> constexpr int kMailWords = (int)(sizeof(T) / 4);  // 64-bit pairs per value
// posts the lane's four values of a strip with sequence number seq
__device__ __forceinline__ void mail_post(unsigned long long* p, unsigned seq, const float (&f)[4]) {
  const unsigned long long hi = (unsigned long long)seq << 32;
  st_mail(p, hi | __float_as_uint(f[0]), hi | __float_as_uint(f[1]));
  st_mail(p + 2, hi | __float_as_uint(f[2]), hi | __float_as_uint(f[3]));
}
__device__ __forceinline__ void mail_post(unsigned long long* p, unsigned seq, const double (&f)[4]) {
  const unsigned long long hi = (unsigned long long)seq << 32;
#pragma unroll
  for (int r = 0; r < 4; r++)
    st_mail(p + 2 * r, hi | (unsigned)__double2loint(f[r]), hi | (unsigned)__double2hiint(f[r]));
}
__device__ __forceinline__ void mail_unpack(const unsigned long long (&w)[4], float (&f)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) f[r] = __uint_as_float((unsigned)w[r]);
}
__device__ __forceinline__ void mail_unpack(const unsigned long long (&w)[8], double (&f)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) f[r] = __hiloint2double((int)(unsigned)w[2 * r + 1], (int)(unsigned)w[2 * r]);
}
// Waits until the lane's four values of a neighbour's strip carry seq and stores them into the halo column.  false: gave up
// (somebody reported an error, or two seconds went by).
template <typename T>
__device__ __forceinline__ bool mail_take(const unsigned long long* p, unsigned seq, T* halo, int* err) {
  constexpr int NQ = 2 * kMailWords<T>;  // 16-byte loads per lane
  unsigned long long w[2 * NQ];
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; spin++) {
    bool ok = true;
#pragma unroll
    for (int q = 0; q < NQ; q++) ld_mail(p + 2 * q, w[2 * q], w[2 * q + 1]);
#pragma unroll
    for (int q = 0; q < 2 * NQ; q++) ok = ok && (unsigned)(w[q] >> 32) == seq;
    if (ok) break;
    if ((spin & 127u) == 127u) {
      if (*(volatile int*)err) return false;
      const unsigned long long t = globaltimer_ns_tile();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) { *err = COL_ERR_TIMEOUT; return false; }
    }
  }
  T f[4];
  mail_unpack(w, f);
  sts4(halo, f);
  return true;
}
// two strips at once: all loads of both are in flight together
template <typename T>
__device__ __forceinline__ bool mail_take2(const unsigned long long* p0, const unsigned long long* p1, unsigned seq, T* halo0, T* halo1,
                                           int* err) {
  constexpr int NQ = 2 * kMailWords<T>;
  unsigned long long w0[2 * NQ], w1[2 * NQ];
  unsigned long long t0 = 0;
  bool ok0 = false, ok1 = false;
  for (unsigned spin = 0;; spin++) {
    if (!ok0) {
#pragma unroll
      for (int q = 0; q < NQ; q++) ld_mail(p0 + 2 * q, w0[2 * q], w0[2 * q + 1]);
    }
    if (!ok1) {
#pragma unroll
      for (int q = 0; q < NQ; q++) ld_mail(p1 + 2 * q, w1[2 * q], w1[2 * q + 1]);
    }
    bool a = true, b = true;
#pragma unroll
    for (int q = 0; q < 2 * NQ; q++) { a = a && (unsigned)(w0[q] >> 32) == seq; b = b && (unsigned)(w1[q] >> 32) == seq; }
    ok0 = a; ok1 = b;
    if (ok0 && ok1) break;
    if ((spin & 127u) == 127u) {
      if (*(volatile int*)err) return false;
      const unsigned long long t = globaltimer_ns_tile();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) { *err = COL_ERR_TIMEOUT; return false; }
    }
  }
  T f[4];
  mail_unpack(w0, f);
  sts4(halo0, f);
  mail_unpack(w1, f);
  sts4(halo1, f);
  return true;
}
// The goal cell (t = 0) and, for nu, its four neighbours (t = 1..4: (+1,0) (-1,0) (0,+1) (0,-1)) are not what the generic
// formula gives (tile_goal_fixup).  The warp that has just stored the strip at buffer offset ob redoes those of them that lie in
// it, one lane per cell, from src; gcell[t] = the cell's offset in a buffer (-1: not in this slab).  A serial version of this
// cost the goal's CTA -- and, through the exchange, every CTA -- 3000 cycles per sweep.
template <typename T, int KIND, bool FAST>
__device__ __forceinline__ void col_goal_patch(const TileK<T>& k, const T* __restrict__ src, T* __restrict__ dst,
                                               const uint32_t* __restrict__ msk, const int* __restrict__ gcell, int RS, int ob, int t) {
  if (t < (KIND == KIND_NU ? 5 : 1)) {
    const int i = gcell[t];
    if (i >= ob && i < ob + TEI) {
      const int m = i - RS;  // the cell bytes have no halo column in front
      const uint32_t b = (msk[m >> 2] >> (8 * (m & 3))) & 0xffu;
      const T c = src[i];
      if (t == 0) {
        // terminal_condition: nu, gap -> 0; logz -> 0 - V(goal), damped like every other state
        if constexpr (kZDom<T, KIND, FAST>) dst[i] = (T)0.5 * k.term + (T)0.5 * c;
        else dst[i] = (KIND == KIND_LOGZ) ? damp_half<T, FAST>(k.term, c) : k.term;
      } else if (!(b & CELL_WALL)) {
        T nb[NA];
        nb[0] = (b & 1u) ? c : src[i + 1];
        nb[1] = (b & 2u) ? c : src[i - 1];
        nb[2] = (b & 4u) ? c : src[i + RS];
        nb[3] = (b & 8u) ? c : src[i - RS];
        const int cidx = k.goal_p + (t == 1) - (t == 2) + k.ld * ((t == 3) - (t == 4));
        dst[i] = backup_nu(k.k, cidx, b, nb);
      }
    }
  }
  __syncwarp();
}

// One of a warp's strips (at most two in fp32, four in Float64) in sweep sw: level sw-1 in src -> level sw in dst.  The warp owns
// the strip for the whole launch, so in fp32 everything that does not change stays in registers: the lane's four values v (its own
// output of the last sweep; they are stored for the neighbours, never reloaded), the cell bytes mk and, for the linear fast
// sweeps, the constant planes pl (a strip of those sweeps is four shared-memory accesses instead of nine: they are bound by the
// shared-memory pipe).  Float64 (KEEPV = false: a strip's arithmetic needs the 128 registers) reloads v and keeps only mk.
// item: offset in a buffer | flags; fix: this lane's seam flags (COL_HAS_UP for lane 0, COL_HAS_DN for lane 31, else 0).
template <typename T> constexpr int kColSlots = sizeof(T) == 8 ? 4 : 2;
template <typename T> constexpr bool kColKeepV = sizeof(T) == 4;
template <typename T, int KIND, bool FAST, bool PL>
__device__ __forceinline__ void col_slot(const TileK<T>& k, const T* __restrict__ src, T* __restrict__ dst, const uint32_t* __restrict__ msk,
                                         const T* __restrict__ auxg, unsigned long long* __restrict__ mine, const int* __restrict__ gcell,
                                         uint32_t item, uint32_t gofs, T (&v)[4], uint32_t mk, T (&pl)[5][4], int wc, int lane4,
                                         uint32_t fix, int RS, unsigned seq, bool post) {
  constexpr int NP = kPlanes<T, KIND, FAST, PL>;
  constexpr int MW = kMailWords<T>;
  const int ob = (int)(item & COL_OFS), o = ob + lane4;
  T L[4], R[4], out[4];
  if constexpr (!kColKeepV<T>) lds4(src + o, v);
  lds4(src + o - RS, L);
  lds4(src + o + RS, R);
  if constexpr (NP == 0 && KIND != KIND_NU) {
#pragma unroll
    for (int q = 0; q < kPlanesGlobal<T, KIND, FAST>; q++) ldg4(auxg + (size_t)q * k.np + gofs + lane4, pl[q]);
  }
  // the cell above a strip's first row / below its last one lives in the neighbouring strip of the same column; outside the
  // bounding box (first / last strip) it is wall and the cell byte blocks the move
  T up = __shfl_up_sync(FULL, v[3], 1), dn = __shfl_down_sync(FULL, v[0], 1);
  if (item & fix & COL_HAS_UP) up = src[o - 1];
  if (item & fix & COL_HAS_DN) dn = src[o + 4];
  strip_compute<T, KIND, FAST>(k, v, L, R, up, dn, mk, pl, out);
  sts4(dst + o, out);
#pragma unroll
  for (int r = 0; r < 4; r++) v[r] = out[r];  // (also what is posted below)
  if (item & COL_GOAL) {  // warp-uniform, at most six strips of the grid
    col_goal_patch<T, KIND, FAST>(k, src, dst, msk, gcell, RS, ob, lane4 >> 2);
    lds4(dst + o, v);
  }
  if (post) {  // the slab's first column is the left neighbour's halo (box 0), its last one the right neighbour's (box 1)
    if (item & COL_EDGE_L) mail_post(mine + (size_t)(o - RS) * MW, seq, v);
    if (item & COL_EDGE_R) mail_post(mine + (size_t)RS * MW + (size_t)(o - wc * RS) * MW, seq, v);
  }
}

// grid = ncta CTAs (one per SM, cooperative launch: all resident) of nwc warps; shared memory: col_smem_bytes().
// sched: the work list [2 shapes: wmax / wmax - 1 columns][nwc warps][G = kColSlots strips] of {item or COL_NONE, -} with item =
// column | strip << 16; a warp's edge-column strip comes first.
template <typename T, int KIND, bool FAST, bool PL>
__global__ void __launch_bounds__(kColMaxThreads<T>, 1)
col_resident_kernel(const TileK<T> k, const ColGeom cg, const uint8_t* __restrict__ maskp, const T* __restrict__ X, T* __restrict__ Y,
                    const T* __restrict__ auxg, const uint2* __restrict__ sched, int nsweeps, int want_res, double* __restrict__ partial,
                    unsigned long long* __restrict__ mail, unsigned seq0, int* __restrict__ err, long long* __restrict__ trace) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ double red[32];
  __shared__ int gcell[5];
  constexpr int NP = kPlanes<T, KIND, FAST, PL>;
  constexpr int MW = kMailWords<T>;
  const int RS = cg.RS, NS = cg.NS;
  // PIQL_TILE_COLRES_TRACE = CTA: lane 0 of every warp of that CTA stamps six points of sweeps 10..13 (timing experiments)
#define COL_STAMP(i) do { if (trace && blockIdx.x == (unsigned)cg.diag_cta && sw >= cg.diag_sw && sw < cg.diag_sw + 4 && lane == 0) \
    trace[((sw - cg.diag_sw) * 32 + warp) * 6 + (i)] = clock64(); } while (0)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
  const int c = blockIdx.x;
  const int wc = cg.wbase + (c < cg.wrem ? 1 : 0);
  const int j0 = cg.c0 + c * cg.wbase + min(c, cg.wrem);
  // two buffers of wmax + 2 columns (a halo column on either side of the slab), the cell bytes of the slab, the work list
  T* bufA = reinterpret_cast<T*>(tile_smem);
  T* bufB = bufA + (size_t)(cg.wmax + 2) * RS;
  uint32_t* msk = reinterpret_cast<uint32_t*>(bufB + (size_t)(cg.wmax + 2) * RS);
  uint4* work = reinterpret_cast<uint4*>(msk + (size_t)cg.wmax * RS / 4);
  const int gr = k.goal_pi - cg.r0, gcol = k.goal_pj - j0;  // the goal relative to the slab (row, column)
  if (threadIdx.x < 5) {  // the goal and its four neighbours: offset in a buffer when the cell is this slab's
    const int t = threadIdx.x, row = gr + (t == 1) - (t == 2), col = gcol + (t == 3) - (t == 4);
    gcell[t] = (col >= 0 && col < wc && row >= 0 && row < RS) ? (col + 1) * RS + row : -1;
  }
  for (int e = threadIdx.x; e < cg.nwc * cg.G; e += blockDim.x) {
    const uint2 in = sched[(size_t)(wc == cg.wmax ? 0 : 1) * cg.nwc * cg.G + e];
    uint32_t it[2] = {in.x, in.y}, gofs[2] = {0u, 0u};
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (it[q] == COL_NONE) { it[q] = COL_END; continue; }
      const int col = (int)(it[q] & 0xffffu), s = (int)(it[q] >> 16);
      uint32_t w = (uint32_t)((col + 1) * RS + 128 * s);
      if (col == 0 && c > 0) w |= COL_EDGE_L;
      if (col == wc - 1 && c < cg.ncta - 1) w |= COL_EDGE_R;
      if (s > 0) w |= COL_HAS_UP;
      if (s < NS - 1) w |= COL_HAS_DN;
      if (col - gcol >= -1 && col - gcol <= 1 && gr - 128 * s >= -1 && gr - 128 * s <= 128) w |= COL_GOAL;
      it[q] = w;
      gofs[q] = (uint32_t)((cg.r0 + 128 * s) + k.ld * (j0 + col));
    }
    work[e] = make_uint4(it[0], it[1], gofs[0], gofs[1]);
  }
  // stage the slab: iterate (with the neighbours' columns of level 0 as halo), cell bytes, constant planes -- once per launch
  const int nitems = wc * NS;
  for (int it = warp; it < (wc + 2) * NS; it += NW) {
    const int col = it / NS - 1, s = it - (col + 1) * NS;
    const int row = 128 * s + 4 * lane, o = (col + 1) * RS + row;
    const size_t g = (size_t)(cg.r0 + row) + (size_t)k.ld * (j0 + col);
    T v[4];
    ldg4(X + g, v);
    sts4(bufA + o, v);
    if (col < 0 || col >= wc) { sts4(bufB + o, v); continue; }  // (both halo copies defined: the first / last slab never receives one)
    msk[(o - RS) >> 2] = __ldg(reinterpret_cast<const uint32_t*>(maskp + g));
  }
  __syncthreads();

  // the warp's strips (edge-column strips first: they are what the neighbours wait for) and their register state
  constexpr int NSL = kColSlots<T>, NKV = kColKeepV<T> ? NSL : 1;
  uint32_t it[NSL], gofs[NSL], mk[NSL];
  T v[NKV][4], pl[NKV][5][4];
  const int lane4 = 4 * lane;
  const uint32_t fix = lane == 0 ? COL_HAS_UP : lane == 31 ? COL_HAS_DN : 0u;
#pragma unroll
  for (int q = 0; q < NSL; q++) {
    const uint4 e = work[warp * cg.G + q];
    it[q] = e.x; gofs[q] = e.z; mk[q] = 0u;
    if (e.x != COL_END) {
      const int o = (int)(e.x & COL_OFS) + lane4;
      mk[q] = msk[(o - RS) >> 2];
      if constexpr (kColKeepV<T>) {
        lds4(bufA + o, v[q]);
#pragma unroll
        for (int r = 0; r < NP; r++) ldg4(auxg + (size_t)r * k.np + e.z + lane4, pl[q][r]);
      }
    }
  }
  T* src = bufA;
  T* dst = bufB;
  const size_t box = (size_t)RS * MW;
  for (int sw = 1; sw <= nsweeps; sw++) {
    int bad = 0;
    const bool exchange = sw < nsweeps && !(cg.diag & 1);  // the last sweep's columns are not needed by anybody
    const unsigned seq = seq0 + (unsigned)sw;
    unsigned long long* mine = mail + (size_t)((sw & 1) * cg.ncta + c) * 2 * box;
    COL_STAMP(0);
#pragma unroll
    for (int q = 0; q < NSL; q++) {
      if (it[q] != COL_END)
        col_slot<T, KIND, FAST, PL>(k, src, dst, msk, auxg, mine, gcell, it[q], gofs[q], v[kColKeepV<T> ? q : 0], mk[q],
                                    pl[kColKeepV<T> ? q : 0], wc, lane4, fix, RS, seq, exchange);
      if (q == 0) COL_STAMP(1);
    }
    COL_STAMP(2);
    // the neighbours' strips of this sweep, posted right after they were computed, into the halo columns of dst: by the warps the
    // work list names (all-interior warps first: they finish ~1000 cycles before the warps that own edge strips).  Tried and dropped (us per nu sweep at 1026^2 against 2.8): four extra warps that do nothing but
    // collect mail (4.0: under this kernel's shared-memory load a global round trip costs 1000-2000 cycles -- a MEMBAR.GPU behind
    // the post takes 2000 -- and their probe-then-fetch needed two); the other warps held back on a named barrier until the
    // edge strips are posted (3.1: the strips finish 800 cycles earlier, their mail arrives no sooner); a fence or a load
    // right behind the post (3.4 / 2.9); two phases with the columns next to the edge in the first (3.3).
    if (exchange) {
      // (the duties -- work list .y: the edge strip whose halo twin this warp fetches -- are read here, not kept in registers)
      const unsigned long long* fromL = mail + ((size_t)((sw & 1) * cg.ncta + c - 1) * 2 + 1) * box;  // left neighbour's last column
      const unsigned long long* fromR = mail + ((size_t)((sw & 1) * cg.ncta + c + 1) * 2 + 0) * box;  // right neighbour's first one
      T* const haloR = dst + (size_t)(wc + 1) * RS;
      if constexpr (NP == 0 && sizeof(T) == 4 && NSL == 2) {
        // two strips' loads in flight together where the registers allow it (a round trip is ~1000 cycles)
        const uint32_t d0 = work[warp * cg.G].y, d1 = work[warp * cg.G + 1].y;
        const int ob0 = (int)(d0 & COL_OFS), ob1 = (int)(d1 & COL_OFS);
        const bool l0 = (d0 & COL_EDGE_L) != 0u, l1 = (d1 & COL_EDGE_L) != 0u;
        const int r0 = (l0 ? ob0 - RS : ob0 - wc * RS) + lane4, r1 = (l1 ? ob1 - RS : ob1 - wc * RS) + lane4;
        const bool has0 = (d0 & (COL_EDGE_L | COL_EDGE_R)) != 0u, has1 = (d1 & (COL_EDGE_L | COL_EDGE_R)) != 0u;
        const unsigned long long* p0 = (l0 ? fromL : fromR) + (size_t)r0 * MW;
        const unsigned long long* p1 = (l1 ? fromL : fromR) + (size_t)r1 * MW;
        T* const h0 = (l0 ? dst : haloR) + r0;
        T* const h1 = (l1 ? dst : haloR) + r1;
        if (has0 && has1) { if (!mail_take2<T>(p0, p1, seq, h0, h1, err)) bad = 1; }
        else if (has0) { if (!mail_take<T>(p0, seq, h0, err)) bad = 1; }
        else if (has1) { if (!mail_take<T>(p1, seq, h1, err)) bad = 1; }
        // (a slab of one column: the same strip faces both neighbours)
        if ((d0 & COL_EDGE_L) && (d0 & COL_EDGE_R)) { if (!mail_take<T>(fromR + (size_t)(ob0 - wc * RS + lane4) * MW, seq, haloR + ob0 - wc * RS + lane4, err)) bad = 1; }
        if ((d1 & COL_EDGE_L) && (d1 & COL_EDGE_R)) { if (!mail_take<T>(fromR + (size_t)(ob1 - wc * RS + lane4) * MW, seq, haloR + ob1 - wc * RS + lane4, err)) bad = 1; }
      } else {
        // (with the constant planes in registers there is none to spare: the owner of an edge strip fetches its halo twin --
        // the strip's own flags -- and nothing is read from the work list here; Float64 follows the list, one strip at a time)
#pragma unroll
        for (int q = 0; q < NSL; q++) {
          const uint32_t d = NP > 0 ? it[q] : work[warp * cg.G + q].y;
          const int ob = (int)(d & COL_OFS);
          if (d & COL_EDGE_L) {
            const int row = ob - RS + lane4;
            if (!mail_take<T>(fromL + (size_t)row * MW, seq, dst + row, err)) bad = 1;
          }
          if (d & COL_EDGE_R) {
            const int row = ob - wc * RS + lane4;
            if (!mail_take<T>(fromR + (size_t)row * MW, seq, haloR + row, err)) bad = 1;
          }
        }
      }
    }
    COL_STAMP(3);
    T* t = src; src = dst; dst = t;
    COL_STAMP(4);
    if (__syncthreads_or(bad)) break;  // block-uniform: a neighbour never answered (the host reports it)
    COL_STAMP(5);
  }
#undef COL_STAMP

  // write the slab back (src: last iterate, dst: the one before); residual of the last sweep
  double res = 0.0;
  bool range = false;
  for (int it = warp; it < nitems; it += NW) {
    const int col = it / NS, s = it - col * NS;
    const int row = 128 * s + 4 * lane, o = (col + 1) * RS + row;
    const size_t g = (size_t)(cg.r0 + row) + (size_t)k.ld * (j0 + col);
    T f[4];
    lds4(src + o, f);
    const uint32_t mk = msk[(o - RS) >> 2];
    if constexpr (kZDom<T, KIND, FAST>) range |= strip_not_normal(f, mk);
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(Y + g) = make_float4((float)f[0], (float)f[1], (float)f[2], (float)f[3]);
    } else {
      *reinterpret_cast<double2*>(Y + g) = make_double2((double)f[0], (double)f[1]);
      *reinterpret_cast<double2*>(Y + g + 2) = make_double2((double)f[2], (double)f[3]);
    }
    if (want_res) {
      T p[4];
      lds4(dst + o, p);
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (!((mk >> (8 * r)) & CELL_WALL))
          res += kZDom<T, KIND, FAST> ? sweep_diff_z((double)f[r], (double)p[r]) : sweep_diff<KIND>((double)f[r], (double)p[r]);
    }
  }
  if (kZDom<T, KIND, FAST> && range) *err = TILE_ERR_RANGE;
  if (want_res) {  // uniform
    const double sum = block_sum(res, red);
    if (threadIdx.x == 0) partial[c] = sum;
  }
}

// ---- packing between the dense H x W arrays of the ABI side and the padded tile layout -------------------
static bool tile_zdom(int precision, int kind, bool fast) { return PIQL_LOGZ_ZDOM && precision != 0 && fast && kind == KIND_LOGZ; }
struct PadGeom {
  int H, W, ld, Wp, kh;  // kh: ghost columns
  int gi;                // ghost rows (= kh for the tiled plans; the column-resident plans align the bounding box instead)
};
// zdom: the padded iterate is kept in the linear domain (fast fp32 log Z): Z = e^(log Z) on the way in, the log on the way out
template <typename T>
__global__ void tile_pack_kernel(PadGeom g, const T* __restrict__ dense, T* __restrict__ padded, int zdom) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.ld * g.Wp) return;
  const int i = p % g.ld - g.gi, j = p / g.ld - g.kh;
  T v = (i >= 0 && i < g.H && j >= 0 && j < g.W) ? dense[i + g.H * j] : (T)0;
  if (zdom) v = (T)exp((double)v);
  padded[p] = v;
}
__global__ void tile_pack_mask_kernel(PadGeom g, const uint8_t* __restrict__ cell, uint8_t* __restrict__ maskp) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.ld * g.Wp) return;
  const int i = p % g.ld - g.gi, j = p / g.ld - g.kh;
  maskp[p] = (i >= 0 && i < g.H && j >= 0 && j < g.W) ? cell[i + g.H * j] : (uint8_t)(CELL_WALL | CELL_BLOCK_MASK);
}
// aux4 [cell][4] + aux1 [cell] -> five planes.  logz: exp(A_a - max A), max A (as prep_kernel stores them); gap: pi_a, D.
// Constants of the linear-domain log Z sweep (logz_cell_lin2): G_a' = 1/2 gamma (w_off sum_a e^(A_a) + dw e^(A_a')) and
// C = 1/2 (1 - gamma) sum_a e^(A_a), in double from exp(A_a - max A) and max A, rounded once.
struct ZPlaneK {
  double gamma, w_off, dw;
};
template <typename T>
__global__ void tile_pack_zplanes_kernel(PadGeom g, ZPlaneK zk, const T* __restrict__ aux4, const T* __restrict__ aux1,
                                         T* __restrict__ auxp) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = g.ld * g.Wp;
  if (p >= np) return;
  const int i = p % g.ld - g.gi, j = p / g.ld - g.kh;
  double o[5] = {0, 0, 0, 0, 0};
  if (i >= 0 && i < g.H && j >= 0 && j < g.W) {
    const int s = i + g.H * j;
    T a[4];
    ldg4(aux4 + 4 * (size_t)s, a);
    const double em = exp((double)aux1[s]);
    double ea[4], sum = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) { ea[q] = (double)a[q] * em; sum += ea[q]; }
#pragma unroll
    for (int q = 0; q < 4; q++) o[q] = 0.5 * zk.gamma * (zk.w_off * sum + zk.dw * ea[q]);
    // (C = 1/2 (1 - gamma) sum = (1 - gamma) / gamma * sum_q o[q] is rebuilt in the sweep: plane 4 stays empty)
  }
#pragma unroll
  for (int q = 0; q < 5; q++) auxp[(size_t)q * np + p] = (T)o[q];
}
template <typename T, int KIND>
__global__ void tile_pack_aux_kernel(PadGeom g, const T* __restrict__ aux4, const T* __restrict__ aux1, T* __restrict__ auxp,
                                     int fold_max) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = g.ld * g.Wp;
  if (p >= np) return;
  const int i = p % g.ld - g.gi, j = p / g.ld - g.kh;
  T o[5] = {0, 0, 0, 0, 0};
  if (i >= 0 && i < g.H && j >= 0 && j < g.W) {
    const int s = i + g.H * j;
    T a[4];
    ldg4(aux4 + 4 * (size_t)s, a);
    {
#pragma unroll
      for (int q = 0; q < 4; q++) o[q] = a[q];
      o[4] = aux1[s];
      if (fold_max) {  // fast fp32 log Z: planes 0..3 hold exp(A_a) = exp(A_a - max A) * exp(max A)
        const T em = (T)exp((double)o[4]);
#pragma unroll
        for (int q = 0; q < 4; q++) o[q] *= em;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; q++) auxp[(size_t)q * np + p] = o[q];
}
template <typename T>
__global__ void tile_unpack_kernel(PadGeom g, const T* __restrict__ padded, T* __restrict__ dense, int zdom) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.H * g.W) return;
  const int i = s % g.H, j = s / g.H;
  const T v = padded[(i + g.gi) + g.ld * (j + g.kh)];
  dense[s] = zdom ? (T)log((double)v) : v;
}

// ---- host side ---------------------------------------------------------------------------------------------
static int tile_planes(int precision, int kind, bool fast) {  // constant planes the linear fast fp32 sweeps can keep in shared memory
  if (!(PIQL_TILE_PLANES && precision != 0 && fast)) return 0;
  return kind == KIND_GAP ? 5 : tile_zdom(precision, kind, fast) ? 4 : 0;
}
static size_t tile_smem_bytes(int EJ, int precision, int planes) {
  return (size_t)EJ * ((2 + planes) * TEI * (precision == 0 ? 8 : 4) + TMB);
}

// Candidate tile shapes (TJ interior columns, nw warps per CTA), cheapest first by a small cost model in SM cycles per
// round of TILE_KH sweeps.  Measured on B200 (fp32 nu): an SM retires one 128-cell warp-strip (~140 instructions) per
// ~61 cycles when ~32 warps are resident, and a warp needs ~1700 cycles per strip however few warps share the SM.
//   multi-launch: waves of c CTAs per SM, each the slower of c*strips*61 and rounds*1700 (strips and rounds summed over
//                 the TILE_KH sweeps of a launch: sweep `sub` computes cols - 2*sub columns), plus a fixed cost per wave
//                 (launch, staging, tail);
//   persistent:   every tile resident at once (tiles <= SMs x CTAs per SM); everybody advances at the pace of the
//                 busiest SM, c*rounds*nw*61 with at least 24 resident warps.
// The model only ranks; tile_begin times the leading candidates on the device and keeps the fastest.
// PIQL_TILE_TJ / PIQL_TILE_NW pin the shape (tests, tuning sweeps).
constexpr size_t TUNE_CANDIDATES = 8;  // shapes of each variant (persistent / multi-launch x depth x planes) tile_begin times on the device
struct PlanCand {
  double cost;
  TilePlan plan;
};
static int tile_kernel_regs(int precision, int kind, bool fast, bool persistent);
// One variant: kh sweeps per round (ghost depth 4 or 8) and `planes` constant planes staged in shared memory (0: read from
// global memory in every sub-sweep).  Costs are per sweep, so variants can be compared.
static std::vector<PlanCand> enumerate_plans(int H, int W, int precision, int kind, bool fast, int num_sms, bool persistent,
                                             int kh = 4, int planes = 0) {
  std::vector<PlanCand> out;
  const int TKH = kh, TTI = TEI - 2 * kh;
  const int tiles_i = (H + TTI - 1) / TTI;
  const char* env_tj = getenv("PIQL_TILE_TJ");
  const char* env_nw = getenv("PIQL_TILE_NW");
  // registers per thread of the instantiation that will run (48 .. 64 in fp32; Float64: 66 for nu with the table-driven
  // exp / log, ~110 for log Z and gap): they, the shared memory of a tile and the 2048-thread limit decide how many CTAs share
  // an SM.  The candidates are timed on the device, so the model only has to offer the right ones.
  const int regs = tile_kernel_regs(precision, kind, fast, persistent);
  const double strip_issue = 61.0, strip_latency = 1700.0, wave_fixed = 8000.0;
  for (int nw = 4; nw <= (precision == 0 ? PIQL_TILE_F64_MAXT : PIQL_TILE_F32_MAXT) / 32; nw++) {
    if (env_nw && atoi(env_nw) != nw) continue;
    for (int TJ = 8; TJ <= 120; TJ++) {
      if (env_tj && atoi(env_tj) != TJ) continue;
      const int EJ = TJ + 2 * TKH, cols = EJ - 2;
      const size_t smem = tile_smem_bytes(EJ, precision, planes);
      if (smem > 200 * 1024) continue;
      int ctas = (int)((227 * 1024) / (smem + 1024));
      ctas = std::min(ctas, std::min(2048 / (32 * nw), 65536 / (regs * 32 * nw)));
      ctas = std::min(ctas, 16);
      if (ctas < 1) continue;
      const int tiles_j = (W + TJ - 1) / TJ;
      const long long tiles = (long long)tiles_i * tiles_j;
      // sweep `sub` of a round computes the cols - 2*sub columns of the trapezoid (tile_subsweep)
      int rounds = 0, strips = 0;
      for (int sub = 0; sub < TKH; sub++) {
        const int cs = std::max(cols - 2 * sub, 0);
        rounds += (cs + nw - 1) / nw;
        strips += cs;
      }
      const long long slots = (long long)num_sms * ctas;
      double cost = 0.0;
      if (persistent) {
        if (tiles > slots) continue;
        const int c = (int)((tiles + num_sms - 1) / num_sms);
        if (c * nw < 24 && !(env_tj || env_nw)) continue;
        cost = (double)c * rounds * nw * strip_issue + 3000.0;
      } else {
        long long left = tiles;
        while (left > 0) {
          const long long now = std::min(left, slots);
          const int c = (int)((now + num_sms - 1) / num_sms);
          cost += wave_fixed + std::max(c * strips * strip_issue, rounds * strip_latency);
          left -= now;
        }
      }
      TilePlan p{};
      p.TJ = TJ; p.EJ = EJ; p.tiles_i = tiles_i; p.tiles_j = tiles_j; p.nw = nw; p.smem = smem; p.ctas_per_sm = ctas;
      p.ld = ((tiles_i * TTI + 2 * TKH + 31) / 32) * 32;
      p.Wp = tiles_j * TJ + 2 * TKH;
      p.np = (long long)p.ld * p.Wp;
      p.kh = kh; p.gi = kh; p.planes = planes;
      out.push_back({cost / kh, p});
    }
  }
  std::stable_sort(out.begin(), out.end(), [](const PlanCand& a, const PlanCand& b) {
    return a.cost < b.cost || (a.cost == b.cost && a.plan.TJ > b.plan.TJ);
  });
  return out;
}
// Column-resident plans (col_resident_kernel): the geometry follows from the grid -- whole columns of the free cells' bounding box,
// split evenly over at most one CTA per SM -- so the candidates differ in warps per CTA and in where the constant planes of the
// linear sweeps live.  Empty when a slab does not fit an SM's shared memory.  PIQL_TILE_COLRES = 0 disables the schedule, = 1
// forces it (tests, tuning sweeps); PIQL_TILE_NW and PIQL_TILE_PLANES_MODE pin a candidate.
static size_t col_smem_bytes(int wmax, int NS, int precision, int planes, int nwc, int G) {
  const size_t col = (size_t)128 * NS, e = precision == 0 ? 8 : 4;
  (void)planes;  // the planes of the linear fast sweeps live in registers
  return 2 * (wmax + 2) * col * e + wmax * col + (size_t)nwc * G * sizeof(uint4);
}
// Work list of a sweep for a slab of wc columns x NS strips on nwc warps: at most nslots strips per warp (two in fp32: they and
// their constants stay in the warp's registers for the whole launch; four in Float64; 0 is returned when the slab has more strips
// than that).  A warp's first strip is an edge-column strip while they last -- the neighbours wait for those -- dealt out round
// robin; every other round goes the other way, so that the warps with an edge strip (which also collect the neighbours' strips at
// the end of the sweep) are the last to get a second one.  item = column | strip << 16; returns the items per warp (nslots).
static int col_schedule(int wc, int NS, int nwc, int nslots, int duties_per_idle, std::vector<uint2>& out) {
  if (wc * NS > nslots * nwc) return 0;
  std::vector<uint32_t> items;
  for (int sidx = 0; sidx < NS; sidx++)
    for (int col : {0, wc - 1})
      if (col == 0 || wc > 1) items.push_back((uint32_t)col | ((uint32_t)sidx << 16));
  for (int sidx = 0; sidx < NS; sidx++)
    for (int col = 1; col < wc - 1; col++) items.push_back((uint32_t)col | ((uint32_t)sidx << 16));
  const int G = nslots;
  out.assign((size_t)nwc * G, make_uint2(COL_NONE, COL_NONE));
  for (size_t q = 0; q < items.size(); q++) {
    const int round = (int)(q / (size_t)nwc), pos = (int)(q % (size_t)nwc);
    const int i = (round & 1) ? nwc - 1 - pos : pos;
    out[(size_t)i * G + round] = make_uint2(items[q], COL_NONE);
  }
  // Who collects the neighbours' strips (.y: the edge strip whose halo twin a warp fetches at the end of the sweep): a warp
  // whose strips are all interior, one each while they last -- the warps that own edge strips are the last to finish (an edge
  // strip also packs and posts) and a fetch is another ~1000 cycles -- and the owner itself for the rest.
  std::vector<int> idle;
  for (int i = nwc - 1; i >= 0; i--) {
    bool edge = false, any = false;
    for (int g = 0; g < G; g++) {
      const uint32_t it = out[(size_t)i * G + g].x;
      if (it == COL_NONE) continue;
      any = true;
      const int col = (int)(it & 0xffffu);
      edge = edge || col == 0 || col == wc - 1;
    }
    if (any && !edge) idle.push_back(i);
  }
  // (duties_per_idle of them per all-interior warp: two where the kernel fetches two strips at once)
  size_t next = 0;
  const size_t room = idle.size() * (size_t)std::min(duties_per_idle, G);
  for (int i = 0; i < nwc; i++)
    for (int g = 0; g < G; g++) {
      const uint32_t it = out[(size_t)i * G + g].x;
      if (it == COL_NONE) continue;
      const int col = (int)(it & 0xffffu);
      if (col != 0 && col != wc - 1) continue;
      if (next < room) { out[(size_t)idle[next % idle.size()] * G + next / idle.size()].y = it; next++; }
      else {
        int slot = g;  // the owner collects: its first free duty slot
        for (int h = 0; h < G; h++) if (out[(size_t)i * G + h].y == COL_NONE) { slot = h; break; }
        out[(size_t)i * G + slot].y = it;
      }
    }
  return G;
}
static std::vector<TilePlan> enumerate_colres(const GridDev& g, int precision, int kind, bool fast, int num_sms) {
  std::vector<TilePlan> out;
  const char* env = getenv("PIQL_TILE_COLRES");
  if ((env && atoi(env) == 0) || num_sms < 1 || g.bi1 < g.bi0 || g.bj1 < g.bj0) return out;
  const int Hb = g.bi1 - g.bi0 + 1, Wb = g.bj1 - g.bj0 + 1;
  const int NS = (Hb + TEI - 1) / TEI;
  const int wmax = (Wb + num_sms - 1) / num_sms, ncta = (Wb + wmax - 1) / wmax;
  if (wmax > 0x7fff || NS > 0xffff) return out;
  const size_t limit = 227 * 1024 - 1024;  // opt-in maximum less the kernel's static shared memory
  const int kh = 4;                         // ghost columns of the padded layout (the schedule itself needs none)
  const int gi = (4 - g.bi0 % 4) % 4, r0 = gi + g.bi0;
  const char* env_nw = getenv("PIQL_TILE_NW");
  const char* env_pm = getenv("PIQL_TILE_PLANES_MODE");
  const int npl = tile_planes(precision, kind, fast);
  for (int pm = (npl > 0 ? 1 : 0); pm >= 0; pm--) {
    if (env_pm && atoi(env_pm) != pm && npl > 0) continue;
    const int planes = pm ? npl : 0;
    // warps: the sweep is wmax * NS strips; the register file allows 28 warps of 72 registers (fp32) or 16 of 128 (Float64)
    // (kColMaxThreads: 28 warps of 72 registers in fp32, 16 of 128 in Float64)
    const int maxw = precision == 0 ? 16 : 28;
    for (int nw = maxw; nw >= 8; nw -= 4) {
      if (env_nw && std::min(atoi(env_nw), maxw) != nw) continue;
      const int nwc = nw;
      std::vector<uint2> tmp;
      const int nslots = precision == 0 ? 4 : 2;  // kColSlots
      const int G = col_schedule(wmax, NS, nwc, nslots, 1, tmp);
      if (G == 0) continue;  // more strips per warp than that
      const size_t smem = col_smem_bytes(wmax, NS, precision, planes, nwc, G);
      if (smem > limit) continue;
      TilePlan p{};
      p.nwc = nwc; p.G = G;
      p.TJ = wmax; p.EJ = wmax; p.kh = kh; p.gi = gi; p.planes = planes; p.tiles_i = 1; p.tiles_j = ncta; p.nw = nw;
      p.ld = ((std::max(gi + g.H, r0 + TEI * NS) + 31) / 32) * 32;
      p.Wp = g.W + 2 * kh;
      p.np = (long long)p.ld * p.Wp;
      p.smem = smem; p.ctas_per_sm = 1;
      p.colres = 1; p.r0 = r0; p.NS = NS; p.c0 = kh + g.bj0; p.ncta = ncta; p.wbase = Wb / ncta; p.wrem = Wb % ncta; p.wmax = wmax;
      out.push_back(p);
    }
  }
  return out;
}
// The variant an untimed plan uses: 4 sweeps per round, planes from global memory, unless the environment pins another
// (PIQL_TILE_KH = 4 | 8, PIQL_TILE_PLANES_MODE = 0 | 1; tests and tuning sweeps).
static void env_variant(int precision, int kind, bool fast, int& kh, int& planes) {
  const char* env_kh = getenv("PIQL_TILE_KH");
  const char* env_pm = getenv("PIQL_TILE_PLANES_MODE");
  kh = env_kh && atoi(env_kh) == 8 ? 8 : 4;
  planes = env_pm && atoi(env_pm) == 1 ? tile_planes(precision, kind, fast) : 0;
}
TilePlan make_tile_plan(int H, int W, int precision, int kind, bool fast, int num_sms) {
  int kh, planes;
  env_variant(precision, kind, fast, kh, planes);
  const std::vector<PlanCand> c = enumerate_plans(H, W, precision, kind, fast, num_sms, false, kh, planes);
  return c.empty() ? TilePlan{} : c[0].plan;
}
// The model's choice: the persistent shape when every tile can be resident at once, else the multi-launch one.
TilePlan choose_tile_plan(int H, int W, int precision, int kind, bool fast, int num_sms, bool* persistent) {
  int kh, planes;
  env_variant(precision, kind, fast, kh, planes);
  const bool allow = getenv("PIQL_TILE_NO_PERSIST") == nullptr && getenv("PIQL_TILE_NO_TMA") == nullptr;
  const std::vector<PlanCand> pp = allow ? enumerate_plans(H, W, precision, kind, fast, num_sms, true, kh, planes) : std::vector<PlanCand>();
  if (persistent) *persistent = !pp.empty();
  return pp.empty() ? make_tile_plan(H, W, precision, kind, fast, num_sms) : pp[0].plan;
}

template <typename T> static TileK<T> make_tile_k(const GridDev& g, const ViState& vs, const TilePlan& pl) {
  TileK<T> k;
  k.ld = pl.ld; k.np = (int)pl.np; k.TJ = pl.TJ; k.EJ = pl.EJ; k.tiles_i = pl.tiles_i; k.tj0 = 0;
  k.kh = pl.kh; k.tti = TEI - 2 * pl.kh;
  k.goal_pi = g.goal % g.H + pl.gi; k.goal_pj = g.goal / g.H + pl.kh; k.goal_p = k.goal_pi + pl.ld * k.goal_pj;
  const double w_on = 1.0 * (1.0 - g.randomness) + g.randomness / NA, w_off = 0.0 * (1.0 - g.randomness) + g.randomness / NA;
  const double a1 = g.gamma * (w_on - w_off);
  const double kc = log(1.0 / NA) + vs.reward_scale * -g.step_cost;
  k.kc = (T)kc; k.kc_lo = (T)(kc - (double)k.kc);
  k.a1 = (T)a1; k.a1_lo = (T)(a1 - (double)k.a1); k.b1 = (T)(g.gamma * w_off); k.a2 = (T)(a1 * 1.4426950408889634);
  k.w_off = (T)w_off; k.dw = (T)(w_on - w_off);
  k.gamma = (T)g.gamma; k.omg = (T)(1.0 - g.gamma);
  k.term = vs.kind == KIND_LOGZ ? (T)(0.0 - vs.v_goal) : (T)0;
  if (sizeof(T) == 4 && tile_zdom(vs.precision, vs.kind, vs.fast)) {
    k.term = (T)exp(0.0 - vs.v_goal);          // Z(goal) = e^(-V(goal))
    k.omg = (T)((1.0 - g.gamma) / g.gamma);    // kappa of logz_cell_lin2
  }
  k.k = make_k<T>(g, vs.reward_scale, vs.beta);
  k.k.H = pl.ld; k.k.n = (int)pl.np; k.k.goal = k.goal_p;
  return k;
}

static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();
  }
  return fn;
}
// 2-D map over a padded array (inner dimension ld, outer Wp), box = 128 x EJ elements, no swizzle
static bool encode_tile_map(void* out, void* base, CUtensorMapDataType dt, size_t esize, const TilePlan& pl, int box_rows = TEI) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)pl.ld, (cuuint64_t)pl.Wp};
  const cuuint64_t gstride[1] = {(cuuint64_t)pl.ld * esize};
  const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)pl.EJ};
  const cuuint32_t estr[2] = {1, 1};
  return fn((CUtensorMap*)out, dt, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 3-D map over the padded constant planes (ld x Wp x 5), box = 128 x EJ x 1
static bool encode_plane_map(void* out, void* base, const TilePlan& pl) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)pl.ld, (cuuint64_t)pl.Wp, 5};
  const cuuint64_t gstride[2] = {(cuuint64_t)pl.ld * 4, (cuuint64_t)pl.np * 4};
  const cuuint32_t box[3] = {(cuuint32_t)TEI, (cuuint32_t)pl.EJ, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void tile_free(TileBuffers& tb) {
  for (void* p : {tb.x[0], tb.x[1], tb.aux, (void*)tb.mask, (void*)tb.partial, (void*)tb.flags, (void*)tb.d_err, (void*)tb.mail, (void*)tb.sched}) if (p) cudaFree(p);
  if (tb.err) cudaFreeHost(tb.err);
  const TileBuffers keep = tb;  // the tuning results and the sharded flag survive a re-allocation
  tb = TileBuffers{};
  tb.ntuned = keep.ntuned; tb.tuned_next = keep.tuned_next;
  for (int i = 0; i < keep.ntuned; i++) {
    tb.tuned_key[i] = keep.tuned_key[i]; tb.tuned_plan[i] = keep.tuned_plan[i]; tb.tuned_persistent[i] = keep.tuned_persistent[i];
  }
  tb.sharded = keep.sharded;
}

template <typename T>
static void pack_t(const GridDev& g, const uint8_t* cell, const ViState& vs, TileBuffers& tb, cudaStream_t st) {
  const PadGeom pg{g.H, g.W, tb.plan.ld, tb.plan.Wp, tb.plan.kh, tb.plan.gi};
  const unsigned nb = nblk(tb.plan.np, 256);
  const int zdom = tile_zdom(vs.precision, vs.kind, vs.fast) ? 1 : 0;
  tile_pack_kernel<T><<<nb, 256, 0, st>>>(pg, (const T*)vs.x[vs.cur], (T*)tb.x[0], zdom);
  tile_pack_kernel<T><<<nb, 256, 0, st>>>(pg, (const T*)vs.x[vs.cur], (T*)tb.x[1], zdom);
  tile_pack_mask_kernel<<<nb, 256, 0, st>>>(pg, cell, tb.mask);
  g_launches += 3;
  if (zdom) {
    const double w_on = 1.0 * (1.0 - g.randomness) + g.randomness / NA, w_off = 0.0 * (1.0 - g.randomness) + g.randomness / NA;
    tile_pack_zplanes_kernel<T><<<nb, 256, 0, st>>>(pg, ZPlaneK{g.gamma, w_off, w_on - w_off}, (const T*)vs.aux4, (const T*)vs.aux1,
                                                    (T*)tb.aux);
    g_launches++;
  } else if (vs.kind == KIND_LOGZ) {
    tile_pack_aux_kernel<T, KIND_LOGZ><<<nb, 256, 0, st>>>(pg, (const T*)vs.aux4, (const T*)vs.aux1, (T*)tb.aux,
                                                           (PIQL_LOGZ_OWN_SHIFT && vs.fast) ? 1 : 0);
    g_launches++;
  } else if (vs.kind == KIND_GAP) {
    tile_pack_aux_kernel<T, KIND_GAP><<<nb, 256, 0, st>>>(pg, (const T*)vs.aux4, (const T*)vs.aux1, (T*)tb.aux, 0);
    g_launches++;
  }
}

// (Re)allocates for the plan, encodes the TMA descriptors and packs vs.x[vs.cur] (+ constants, masks) into the padded layout.
static cudaError_t tile_setup(const GridDev& g, const uint8_t* cell, const ViState& vs, TileBuffers& tb, const TilePlan& pl,
                              bool persist, cudaStream_t st) {
  const size_t e = vs.precision == 0 ? 8 : 4;
  const size_t need_x = e * (size_t)pl.np, need_aux = vs.kind == KIND_NU ? 0 : 5 * need_x;
  const long long ntiles = (long long)pl.tiles_i * pl.tiles_j;
  if (need_x > tb.cap_x || need_aux > tb.cap_aux || (size_t)pl.np > tb.cap_mask || ntiles > tb.cap_partial ||
      ntiles > tb.cap_flags) {
    cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) return err;
    tile_free(tb);
    // sized for double and all five planes so that later solves of any kind / precision reuse them (+ slack for re-tiling)
    const size_t bx = 8 * (size_t)pl.np + 8 * (size_t)pl.np / 8;
    const long long nt = ntiles * 2 + 64;
    if ((err = cudaMalloc(&tb.x[0], bx)) != cudaSuccess) return err;
    if ((err = cudaMalloc(&tb.x[1], bx)) != cudaSuccess) return err;
    if ((err = cudaMalloc(&tb.aux, 5 * bx)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&tb.mask, bx / 8)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&tb.partial, sizeof(double) * (size_t)nt)) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&tb.flags, sizeof(unsigned) * (size_t)nt)) != cudaSuccess) return err;
    if ((err = cudaMallocHost((void**)&tb.err, sizeof(int))) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**)&tb.d_err, sizeof(int))) != cudaSuccess) return err;
    if ((err = cudaMemset(tb.d_err, 0, sizeof(int))) != cudaSuccess) return err;
    *tb.err = 0;
    tb.cap_x = bx; tb.cap_aux = 5 * bx; tb.cap_mask = bx / 8; tb.cap_partial = nt; tb.cap_flags = nt;
  }
  if (pl.colres) {
    // mailboxes of the per-sweep column exchange: [parity][CTA][side][row] pairs of {32-bit word, sequence number}
    const size_t need_mail = 2 * (size_t)pl.ncta * 2 * TEI * pl.NS * (e / 4) * sizeof(unsigned long long);
    if (need_mail > tb.cap_mail) {
      cudaError_t err = cudaStreamSynchronize(st);
      if (err != cudaSuccess) return err;
      if (tb.mail) cudaFree(tb.mail);
      tb.mail = nullptr; tb.cap_mail = 0;
      if ((err = cudaMalloc((void**)&tb.mail, need_mail)) != cudaSuccess) return err;
      if ((err = cudaMemset(tb.mail, 0, need_mail)) != cudaSuccess) return err;
      tb.cap_mail = need_mail;
      tb.col_seq = 0;
    }
  }
  if (pl.colres) {
    // the work list of both slab shapes (wmax / wmax - 1 columns)
    std::vector<uint2> a, b, both;
    // (two duties per all-interior warp where the kernel fetches two strips at once: fp32 without planes in registers)
    const char* env_dp = getenv("PIQL_TILE_COLRES_DUTIES");
    const int dp = env_dp ? atoi(env_dp) : (vs.precision != 0 && pl.planes == 0) ? 2 : 1;
    const int Ga = col_schedule(pl.wmax, pl.NS, pl.nwc, pl.G, dp, a), Gb = pl.wmax > 1 ? col_schedule(pl.wmax - 1, pl.NS, pl.nwc, pl.G, dp, b) : 1;
    if (Ga != pl.G || (pl.wmax > 1 && Gb != pl.G)) return cudaErrorInvalidValue;
    both.assign(2 * (size_t)pl.nwc * pl.G, make_uint2(COL_NONE, COL_NONE));
    for (int i = 0; i < pl.nwc; i++) {
      for (int g = 0; g < Ga; g++) both[(size_t)i * pl.G + g] = a[(size_t)i * Ga + g];
      if (pl.wmax > 1) for (int g = 0; g < Gb; g++) both[((size_t)pl.nwc + i) * pl.G + g] = b[(size_t)i * Gb + g];
    }
    const size_t bytes = both.size() * sizeof(uint2);
    cudaError_t err = cudaStreamSynchronize(st);  // an earlier launch may still read the list
    if (err != cudaSuccess) return err;
    if (bytes > tb.cap_sched) {
      if (tb.sched) cudaFree(tb.sched);
      tb.sched = nullptr; tb.cap_sched = 0;
      if ((err = cudaMalloc((void**)&tb.sched, bytes)) != cudaSuccess) return err;
      tb.cap_sched = bytes;
    }
    if ((err = cudaMemcpy(tb.sched, both.data(), bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return err;
  }
  tb.plan = pl;
  tb.cur = 0;
  // a new solve starts with a clean error word (timed-out waits, exponent-domain range)
  if (cudaMemsetAsync(tb.d_err, 0, sizeof(int), st) != cudaSuccess ||
      cudaMemcpyAsync(tb.err, tb.d_err, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
    return cudaGetLastError();
  static_assert(sizeof(CUtensorMap) == 128, "descriptor size");
  const CUtensorMapDataType dt = vs.precision == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  tb.use_tma = !pl.colres && getenv("PIQL_TILE_NO_TMA") == nullptr && encode_tile_map(tb.tmx[0], tb.x[0], dt, e, pl) &&
               encode_tile_map(tb.tmx[1], tb.x[1], dt, e, pl) &&
               encode_tile_map(tb.tmm, tb.mask, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, pl, TMB) &&
               (pl.planes == 0 || encode_plane_map(tb.tmp, tb.aux, pl));
  tb.persistent = persist && tb.use_tma;
  if (vs.precision == 0) pack_t<double>(g, cell, vs, tb, st);
  else pack_t<float>(g, cell, vs, tb, st);
  tb.active = true;
  return cudaGetLastError();
}

cudaError_t tile_begin(const GridDev& g, const uint8_t* cell, const ViState& vs, TileBuffers& tb, int num_sms, cudaStream_t st) {
  const long long key = (((long long)g.H * 65537 + g.W) * 4 + vs.precision * 2 + (vs.fast ? 1 : 0)) * 4 + vs.kind;
  if (tb.sharded) {  // every rank must use the same shape: the model's multi-launch choice, nothing timed
    const TilePlan pl = make_tile_plan(g.H, g.W, vs.precision, vs.kind, vs.fast, num_sms);
    return tile_setup(g, cell, vs, tb, pl, false, st);
  }
  const std::vector<TilePlan> cr = enumerate_colres(g, vs.precision, vs.kind, vs.fast, num_sms);
  if (!cr.empty() && getenv("PIQL_TILE_COLRES") != nullptr) return tile_setup(g, cell, vs, tb, cr[0], false, st);  // forced
  const bool pinned = getenv("PIQL_TILE_TJ") || getenv("PIQL_TILE_NW") || getenv("PIQL_TILE_NO_TUNE");
  if (!pinned)
    for (int i = 0; i < tb.ntuned; i++)
      if (tb.tuned_key[i] == key) return tile_setup(g, cell, vs, tb, tb.tuned_plan[i], tb.tuned_persistent[i], st);
  bool persist = false;
  const TilePlan model = choose_tile_plan(g.H, g.W, vs.precision, vs.kind, vs.fast, num_sms, &persist);
  if (pinned) return tile_setup(g, cell, vs, tb, model, persist, st);
  // time the leading candidates of every variant on the device (a few launches each) and keep the fastest for this problem:
  // persistent / one launch per round, 4 or 8 sweeps per round (PIQL_TILE_KH pins it), constant planes in shared memory or
  // read from global memory (the two linear fast fp32 sweeps; PIQL_TILE_PLANES_MODE = 0 / 1 pins it)
  struct Try { TilePlan plan; bool persist; };
  std::vector<Try> tries;
  const bool allow_p = getenv("PIQL_TILE_NO_PERSIST") == nullptr && getenv("PIQL_TILE_NO_TMA") == nullptr;
  const char* env_kh = getenv("PIQL_TILE_KH");
  const char* env_pm = getenv("PIQL_TILE_PLANES_MODE");
  const int npl = tile_planes(vs.precision, vs.kind, vs.fast);
  for (int kh = 4; kh <= TILE_KH_MAX; kh += 4) {
    if (env_kh && atoi(env_kh) != kh) continue;
    for (int pm = 0; pm < (npl > 0 ? 2 : 1); pm++) {
      if (env_pm && atoi(env_pm) != pm) continue;
      const int planes = pm ? npl : 0;
      if (allow_p) {
        const std::vector<PlanCand> pp = enumerate_plans(g.H, g.W, vs.precision, vs.kind, vs.fast, num_sms, true, kh, planes);
        for (size_t i = 0; i < pp.size() && i < TUNE_CANDIDATES; i++) tries.push_back({pp[i].plan, true});
      }
      const std::vector<PlanCand> mm = enumerate_plans(g.H, g.W, vs.precision, vs.kind, vs.fast, num_sms, false, kh, planes);
      for (size_t i = 0; i < mm.size() && i < TUNE_CANDIDATES; i++) tries.push_back({mm[i].plan, false});
    }
  }
  for (const TilePlan& p : cr) tries.push_back({p, false});  // every cell once per sweep, slabs resident in shared memory
  cudaEvent_t e0, e1;
  cudaError_t err = cudaEventCreate(&e0);
  if (err != cudaSuccess) return err;
  if ((err = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return err; }
  Try chosen{model, persist};
  // Short probes (small grids: 0.1 ms) are noisy and are repeated; once a probe lasts more than half a millisecond one
  // measurement per shape is enough, and all a first solve should pay (tuning stays near a tenth of a solve to tolerance).
  int reps = 3;
  auto time_try = [&](const Try& t, long long probe, float& ms, bool& persisted) -> cudaError_t {
    cudaError_t e = tile_setup(g, cell, vs, tb, t.plan, t.persist, st);
    if (e != cudaSuccess) return e;
    if ((e = tile_sweeps(g, vs, tb, 2 * t.plan.kh, false, st)) != cudaSuccess) return e;
    ms = 1e30f;
    for (int rep = 0; rep < reps; rep++) {
      cudaEventRecord(e0, st);
      if ((e = tile_sweeps(g, vs, tb, probe, false, st)) != cudaSuccess) return e;
      cudaEventRecord(e1, st);
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
      float m1 = 0.f;
      cudaEventElapsedTime(&m1, e0, e1);
      ms = std::min(ms, m1);
      if (m1 > 0.5f) reps = 1;
    }
    persisted = tb.persistent;
    return (tb.err && *tb.err) ? cudaErrorLaunchTimeout : cudaSuccess;
  };
  // first every candidate with a short probe, then the three fastest and the fastest column-resident one again with a probe
  // of 128 sweeps -- the length of a solver's batches: a whole batch in one launch (column-resident) and one launch per round
  // differ by a few per cent, which 32 sweeps do not resolve
  std::vector<std::pair<float, size_t>> rank;
  for (size_t i = 0; i < tries.size() && err == cudaSuccess; i++) {
    float ms = 0.f;
    bool p = false;
    if ((err = time_try(tries[i], 32, ms, p)) == cudaSuccess) { tries[i].persist = p; rank.push_back({ms, i}); }
    else if (err == cudaErrorCooperativeLaunchTooLarge && tries[i].plan.colres) {
      // the slabs cannot all be resident (SMs taken by another context): the schedule is simply not a candidate here
      cudaGetLastError();
      err = cudaSuccess;
    }
  }
  if (err == cudaSuccess && !rank.empty()) {
    std::stable_sort(rank.begin(), rank.end());
    std::vector<size_t> finals;
    for (size_t q = 0; q < rank.size() && q < 3; q++) finals.push_back(rank[q].second);
    for (size_t q = 0; q < rank.size(); q++)
      if (tries[rank[q].second].plan.colres) {
        if (std::find(finals.begin(), finals.end(), rank[q].second) == finals.end()) finals.push_back(rank[q].second);
        break;
      }
    chosen = tries[finals[0]];
    if (finals.size() > 1 && rank[0].first < 2.0f) {  // (a long probe of a slow grid would cost more than it can gain)
      float best = 1e30f;
      for (size_t i : finals) {
        float ms = 0.f;
        bool p = false;
        if ((err = time_try(tries[i], 128, ms, p)) != cudaSuccess) break;
        if (ms < best) { best = ms; chosen = tries[i]; chosen.persist = p; }
      }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (err != cudaSuccess) return err;
  const int slot = tb.ntuned < TileBuffers::TUNED_MAX ? tb.ntuned++ : (tb.tuned_next++ % TileBuffers::TUNED_MAX);
  tb.tuned_key[slot] = key; tb.tuned_plan[slot] = chosen.plan; tb.tuned_persistent[slot] = chosen.persist;
  return tile_setup(g, cell, vs, tb, chosen.plan, chosen.persist, st);  // re-pack: the probes advanced the iterate
}

template <typename T, int KIND, bool FAST, bool PL>
static cudaError_t launch_tile_t(const GridDev& g, const ViState& vs, TileBuffers& tb, int nsub, bool want_res, cudaStream_t st) {
  static bool configured[64] = {};  // shared-memory opt-in, once per instantiation and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tile_sweep_kernel<T, KIND, FAST, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  TileK<T> k = make_tile_k<T>(g, vs, tb.plan);
  k.tj0 = tb.sharded ? tb.tj0 : 0;
  const unsigned grid = (unsigned)(tb.plan.tiles_i * (tb.sharded ? tb.tj1 - tb.tj0 : tb.plan.tiles_j));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * tb.plan.nw); cfg.dynamicSmemBytes = tb.plan.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CUtensorMap tmX, tmM, tmP;
  memcpy(&tmX, tb.tmx[tb.cur], sizeof(tmX));
  memcpy(&tmM, tb.tmm, sizeof(tmM));
  memcpy(&tmP, tb.tmp, sizeof(tmP));
  TileHalo halo = tb.sharded ? tb.halo : TileHalo();
  halo.err = tb.d_err;  // also where a linear-domain sweep reports that it left the float range
  cudaError_t le = cudaLaunchKernelEx(&cfg, tile_sweep_kernel<T, KIND, FAST, PL>, k, tmX, tmM, tmP, tb.use_tma ? 1 : 0,
                                      (const uint8_t*)tb.mask, (const T*)tb.x[tb.cur],
                                      (T*)tb.x[tb.cur ^ 1], (const T*)tb.aux, nsub, want_res ? 1 : 0, tb.partial, halo);
  if (le != cudaSuccess) return le;
  g_launches++;
  tb.cur ^= 1;
  return cudaSuccess;
}

// All rounds of nsweeps sweeps in one cooperative launch (tb.persistent).  cudaErrorCooperativeLaunchTooLarge means the
// occupancy estimate of the plan was optimistic for this instantiation: the caller falls back to one launch per round.
template <typename T, int KIND, bool FAST, bool PL>
static cudaError_t launch_persist_t(const GridDev& g, const ViState& vs, TileBuffers& tb, long long nsweeps, bool want_res,
                                    cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tile_persist_kernel<T, KIND, FAST, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const TileK<T> k = make_tile_k<T>(g, vs, tb.plan);
  const unsigned grid = (unsigned)(tb.plan.tiles_i * tb.plan.tiles_j);
  const int TKH = tb.plan.kh;
  const int nrounds = (int)((nsweeps + TKH - 1) / TKH), nsub_last = (int)(nsweeps - (long long)(nrounds - 1) * TKH);
  cudaError_t e = cudaMemsetAsync(tb.flags, 0, sizeof(unsigned) * grid, st);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * tb.plan.nw); cfg.dynamicSmemBytes = tb.plan.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CUtensorMap tmA, tmB, tmM, tmP;
  memcpy(&tmA, tb.tmx[tb.cur], sizeof(tmA));
  memcpy(&tmB, tb.tmx[tb.cur ^ 1], sizeof(tmB));
  memcpy(&tmM, tb.tmm, sizeof(tmM));
  memcpy(&tmP, tb.tmp, sizeof(tmP));
  e = cudaLaunchKernelEx(&cfg, tile_persist_kernel<T, KIND, FAST, PL>, k, tmA, tmB, tmM, tmP, (T*)tb.x[tb.cur], (T*)tb.x[tb.cur ^ 1],
                         (const T*)tb.aux, nrounds, nsub_last, want_res ? 1 : 0, tb.partial, tb.flags, tb.plan.tiles_j, tb.d_err);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(tb.err, tb.d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return e;
  g_launches++;
  tb.cur = (tb.cur + nrounds) & 1;
  return cudaSuccess;
}
// The column-resident schedule: all nsweeps sweeps of the batch in one cooperative launch of one CTA per slab.
template <typename T, int KIND, bool FAST, bool PL>
static cudaError_t launch_colres_t(const GridDev& g, const ViState& vs, TileBuffers& tb, int nsweeps, bool want_res, cudaStream_t st) {
  static size_t configured[64] = {};  // shared-memory opt-in per instantiation and device (grows with the plan)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || configured[dev] < tb.plan.smem) {
    cudaError_t e = cudaFuncSetAttribute(col_resident_kernel<T, KIND, FAST, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb.plan.smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = tb.plan.smem;
  }
  const TilePlan& pl = tb.plan;
  const TileK<T> k = make_tile_k<T>(g, vs, pl);
  const char* env_diag = getenv("PIQL_TILE_COLRES_DIAG");
  const char* env_trace = getenv("PIQL_TILE_COLRES_TRACE");
  const ColGeom cg{pl.r0, pl.NS, TEI * pl.NS, pl.c0, pl.wbase, pl.wrem, pl.wmax, pl.ncta, pl.nwc, pl.G, env_diag ? atoi(env_diag) : 0,
                   env_trace ? atoi(env_trace) : 0, getenv("PIQL_TILE_COLRES_TRACE_SW") ? atoi(getenv("PIQL_TILE_COLRES_TRACE_SW")) : 10};
  long long* d_trace = nullptr;
  if (env_trace && nsweeps >= cg.diag_sw + 4) {
    if (cudaMalloc((void**)&d_trace, sizeof(long long) * 4 * 32 * 6) != cudaSuccess) return cudaGetLastError();
    cudaMemsetAsync(d_trace, 0, sizeof(long long) * 4 * 32 * 6, st);
  }
  cudaError_t e;
  if (tb.col_seq > 0xC0000000u) {  // sequence numbers wrap: start over on clean mailboxes
    if ((e = cudaMemsetAsync(tb.mail, 0, tb.cap_mail, st)) != cudaSuccess) return e;
    tb.col_seq = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl.ncta); cfg.blockDim = dim3(32 * pl.nw); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, col_resident_kernel<T, KIND, FAST, PL>, k, cg, (const uint8_t*)tb.mask, (const T*)tb.x[tb.cur],
                         (T*)tb.x[tb.cur ^ 1], (const T*)tb.aux, (const uint2*)tb.sched, nsweeps, want_res ? 1 : 0, tb.partial, tb.mail,
                         tb.col_seq, tb.d_err, d_trace);
  if (e != cudaSuccess) return e;
  if (d_trace) {  // timing experiment: print the stamps of sweeps 10..13 relative to the first one (cycles)
    std::vector<long long> h(4 * 32 * 6);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    cudaFree(d_trace);
    long long t0 = 0;
    for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
    for (int q = 0; q < 4; q++)
      for (int w = 0; w < pl.nw; w++) {
        fprintf(stderr, "trace sweep %d warp %2d:", cg.diag_sw + q, w);
        for (int i = 0; i < 6; i++) fprintf(stderr, " %7lld", h[(q * 32 + w) * 6 + i] ? h[(q * 32 + w) * 6 + i] - t0 : -1);
        fprintf(stderr, "\n");
      }
  }
  g_launches++;
  tb.col_seq += (unsigned)nsweeps;
  tb.cur ^= 1;
  return cudaSuccess;
}
template <typename T, bool FAST>
static cudaError_t launch_colres_k(const GridDev& g, const ViState& vs, TileBuffers& tb, int nsweeps, bool want_res, cudaStream_t st) {
  const bool pl = tb.plan.planes > 0;
  switch (vs.kind) {
    case KIND_NU: return launch_colres_t<T, KIND_NU, FAST, false>(g, vs, tb, nsweeps, want_res, st);
    case KIND_LOGZ:
      if constexpr (FAST) { if (pl) return launch_colres_t<T, KIND_LOGZ, FAST, true>(g, vs, tb, nsweeps, want_res, st); }
      return launch_colres_t<T, KIND_LOGZ, FAST, false>(g, vs, tb, nsweeps, want_res, st);
    default:
      if constexpr (FAST) { if (pl) return launch_colres_t<T, KIND_GAP, FAST, true>(g, vs, tb, nsweeps, want_res, st); }
      return launch_colres_t<T, KIND_GAP, FAST, false>(g, vs, tb, nsweeps, want_res, st);
  }
}
// PL = true exists for the two linear fast fp32 sweeps only (the plan asks for it through TilePlan::planes)
template <typename T, bool FAST>
static cudaError_t launch_persist_k(const GridDev& g, const ViState& vs, TileBuffers& tb, long long nsweeps, bool want_res,
                                    cudaStream_t st) {
  const bool pl = tb.plan.planes > 0;
  switch (vs.kind) {
    case KIND_NU: return launch_persist_t<T, KIND_NU, FAST, false>(g, vs, tb, nsweeps, want_res, st);
    case KIND_LOGZ:
      if constexpr (FAST) { if (pl) return launch_persist_t<T, KIND_LOGZ, FAST, true>(g, vs, tb, nsweeps, want_res, st); }
      return launch_persist_t<T, KIND_LOGZ, FAST, false>(g, vs, tb, nsweeps, want_res, st);
    default:
      if constexpr (FAST) { if (pl) return launch_persist_t<T, KIND_GAP, FAST, true>(g, vs, tb, nsweeps, want_res, st); }
      return launch_persist_t<T, KIND_GAP, FAST, false>(g, vs, tb, nsweeps, want_res, st);
  }
}

template <typename T, bool FAST>
static cudaError_t launch_tile_k(const GridDev& g, const ViState& vs, TileBuffers& tb, int nsub, bool want_res, cudaStream_t st) {
  const bool pl = tb.plan.planes > 0;
  switch (vs.kind) {
    case KIND_NU: return launch_tile_t<T, KIND_NU, FAST, false>(g, vs, tb, nsub, want_res, st);
    case KIND_LOGZ:
      if constexpr (FAST) { if (pl) return launch_tile_t<T, KIND_LOGZ, FAST, true>(g, vs, tb, nsub, want_res, st); }
      return launch_tile_t<T, KIND_LOGZ, FAST, false>(g, vs, tb, nsub, want_res, st);
    default:
      if constexpr (FAST) { if (pl) return launch_tile_t<T, KIND_GAP, FAST, true>(g, vs, tb, nsub, want_res, st); }
      return launch_tile_t<T, KIND_GAP, FAST, false>(g, vs, tb, nsub, want_res, st);
  }
}

// Registers per thread of the kernel a (precision, kind, fast) solve runs, asked of the runtime once per instantiation;
// the __launch_bounds__ cap when the query fails.
template <typename T, int KIND, bool FAST> static int regs_of(bool persistent) {
  static int cache[2] = {0, 0};
  int& c = cache[persistent ? 1 : 0];
  if (c == 0) {
    cudaFuncAttributes a;
    cudaError_t e = persistent ? cudaFuncGetAttributes(&a, tile_persist_kernel<T, KIND, FAST, false>)
                               : cudaFuncGetAttributes(&a, tile_sweep_kernel<T, KIND, FAST, false>);
    if (e == cudaSuccess && a.numRegs > 0) c = ((a.numRegs + 7) / 8) * 8;  // allocation granularity: 8 per thread
    else { cudaGetLastError(); c = sizeof(T) == 4 ? 64 : 128; }
  }
  return c;
}
template <typename T, bool FAST> static int regs_of_kind(int kind, bool persistent) {
  switch (kind) {
    case KIND_NU: return regs_of<T, KIND_NU, FAST>(persistent);
    case KIND_LOGZ: return regs_of<T, KIND_LOGZ, FAST>(persistent);
    default: return regs_of<T, KIND_GAP, FAST>(persistent);
  }
}
static int tile_kernel_regs(int precision, int kind, bool fast, bool persistent) {
  if (precision == 0) return regs_of_kind<double, false>(kind, persistent);
  return fast ? regs_of_kind<float, true>(kind, persistent) : regs_of_kind<float, false>(kind, persistent);
}

// nsweeps Jacobi sweeps in ceil(nsweeps / TILE_KH) launches; the residual of the last sweep lands in vs.residual.
cudaError_t tile_sweeps(const GridDev& g, const ViState& vs, TileBuffers& tb, long long nsweeps, bool want_res, cudaStream_t st) {
  const int TKH = tb.plan.kh;
  if (tb.plan.colres) {
    if (tb.err && *tb.err == COL_ERR_TIMEOUT) return cudaErrorLaunchTimeout;  // an earlier launch gave up waiting for a neighbour
    while (nsweeps > 0) {
      const int batch = (int)std::min<long long>(nsweeps, 1 << 24);
      nsweeps -= batch;
      const bool res = want_res && nsweeps == 0;
      cudaError_t e;
      if (vs.precision == 0) e = launch_colres_k<double, false>(g, vs, tb, batch, res, st);
      else if (vs.fast) e = launch_colres_k<float, true>(g, vs, tb, batch, res, st);
      else e = launch_colres_k<float, false>(g, vs, tb, batch, res, st);
      if (e != cudaSuccess) return e;
      if (res) launch_residual_final(tb.partial, tb.plan.ncta, vs.residual, st);
    }
    cudaError_t e = cudaMemcpyAsync(tb.err, tb.d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
    return e != cudaSuccess ? e : cudaGetLastError();
  }
  if (tb.persistent && nsweeps > TKH) {
    if (tb.err && *tb.err) return cudaErrorLaunchTimeout;  // an earlier persistent launch gave up waiting for a neighbour
    cudaError_t e;
    if (vs.precision == 0) e = launch_persist_k<double, false>(g, vs, tb, nsweeps, want_res, st);
    else if (vs.fast) e = launch_persist_k<float, true>(g, vs, tb, nsweeps, want_res, st);
    else e = launch_persist_k<float, false>(g, vs, tb, nsweeps, want_res, st);
    if (e == cudaSuccess) {
      if (want_res) launch_residual_final(tb.partial, tb.plan.tiles_i * tb.plan.tiles_j, vs.residual, st);
      return cudaGetLastError();
    }
    if (e != cudaErrorCooperativeLaunchTooLarge) return e;
    cudaGetLastError();
    tb.persistent = false;  // not all tiles fit at once after all: one launch per round from here on
  }
  while (nsweeps > 0) {
    const int nsub = (int)std::min<long long>(nsweeps, TKH);
    nsweeps -= nsub;
    const bool res = want_res && nsweeps == 0;
    cudaError_t e;
    if (vs.precision == 0) e = launch_tile_k<double, false>(g, vs, tb, nsub, res, st);
    else if (vs.fast) e = launch_tile_k<float, true>(g, vs, tb, nsub, res, st);
    else e = launch_tile_k<float, false>(g, vs, tb, nsub, res, st);
    if (e != cudaSuccess) return e;
    if (res) launch_residual_final(tb.partial, tb.plan.tiles_i * tb.plan.tiles_j, vs.residual, st);
  }
  if (tile_zdom(vs.precision, vs.kind, vs.fast)) {  // the range word travels with the batch (read after the next sync)
    cudaError_t e = cudaMemcpyAsync(tb.err, tb.d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

// Copies the current iterate back into the dense array vs.x[vs.cur].
void tile_unpack(const GridDev& g, const ViState& vs, const TileBuffers& tb, cudaStream_t st) {
  const PadGeom pg{g.H, g.W, tb.plan.ld, tb.plan.Wp, tb.plan.kh, tb.plan.gi};
  const int zdom = tile_zdom(vs.precision, vs.kind, vs.fast) ? 1 : 0;
  if (vs.precision == 0) tile_unpack_kernel<double><<<nblk(g.n, 256), 256, 0, st>>>(pg, (const double*)tb.x[tb.cur], (double*)vs.x[vs.cur], 0);
  else tile_unpack_kernel<float><<<nblk(g.n, 256), 256, 0, st>>>(pg, (const float*)tb.x[tb.cur], (float*)vs.x[vs.cur], zdom);
  g_launches++;
}

// ---- multi-GPU soft VI: column slabs of tiles, halo columns exchanged over NVLink peer memory ---------------------------
// Every rank holds the whole padded array (the layout is replicated) but sweeps only the tile columns [tj0, tj1) it owns.
// The exchange rides in the tile kernel (TileHalo): the CTAs of the slab's first / last tile column store the first / last
// TILE_KH interior columns of the new iterate also into the same places of the left / right neighbour's array (its halo)
// with peer stores, and the last of them raises the neighbour's counter; in the next round those CTAs -- and only those --
// wait for the neighbours' counters before they stage.  Receiving round r-1 from a neighbour also means the neighbour's
// edge tiles have finished reading the buffer that round r's push overwrites (interior tiles never read the halo).  Jacobi + tiling-independent values make
// the result bit-identical to the single-GPU solve for any number of ranks.
cudaError_t tile_sweeps_sharded(const GridDev& g, const ViState& vs, TileBuffers& tb, ViPeer& vp, long long nsweeps, bool want_res,
                                cudaStream_t st) {
  const bool has_l = vp.rank > 0, has_r = vp.rank < vp.world - 1;
  const int TKH = tb.plan.kh;
  while (nsweeps > 0) {
    const int nsub = (int)std::min<long long>(nsweeps, TKH);
    nsweeps -= nsub;
    const bool res = want_res && nsweeps == 0;
    // one launch per round, as on one GPU: its edge CTAs wait for the neighbours' previous round and push this one
    const int out = tb.cur ^ 1;
    TileHalo h;
    h.in_left = has_l ? vp.own_flags + 0 : nullptr;
    h.in_right = has_r ? vp.own_flags + 1 : nullptr;
    h.wait_epoch = (unsigned)vp.round;
    h.out_left = has_l ? vp.x[out][vp.rank - 1] : nullptr;
    h.out_right = has_r ? vp.x[out][vp.rank + 1] : nullptr;
    // my first columns are the left neighbour's right halo (its counter "from right"), and vice versa
    h.flag_left = has_l ? vp.flags[vp.rank - 1] + 1 : nullptr;
    h.flag_right = has_r ? vp.flags[vp.rank + 1] + 0 : nullptr;
    h.done = vp.done;
    h.post_epoch = (unsigned)(vp.round + 1);
    h.tj_first = tb.tj0; h.tj_last = tb.tj1 - 1;
    h.err = tb.d_err;
    h.wait_stats = vp.wait_stats;
    tb.halo = h;
    cudaError_t e;
    if (vs.precision == 0) e = launch_tile_k<double, false>(g, vs, tb, nsub, res, st);
    else if (vs.fast) e = launch_tile_k<float, true>(g, vs, tb, nsub, res, st);
    else e = launch_tile_k<float, false>(g, vs, tb, nsub, res, st);
    if (e != cudaSuccess) return e;
    vp.round++;
    if (res) launch_residual_final(tb.partial, tb.plan.tiles_i * (tb.tj1 - tb.tj0), vs.residual, st);
  }
  cudaError_t e = cudaMemcpyAsync(tb.err, tb.d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace piql
