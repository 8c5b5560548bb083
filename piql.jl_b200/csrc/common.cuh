// common.cuh -- shared device helpers for the PIQL B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace piql {

constexpr int NA = 4;                  // actions of the 2-D gridworld (step_choices, GridWorld.jl:20-22)
constexpr uint8_t CELL_BLOCK_MASK = 0x0f;  // bit a: take_step(s,a) is blocked by a wall
constexpr uint8_t CELL_WALL = 0x10;
constexpr uint8_t CELL_GOAL = 0x20;
constexpr unsigned FULL = 0xffffffffu;

// Problem constants, passed by value to every kernel.
struct GridDev {
  int H, W;
  int n;       // H*W
  int goal;    // linear index of the goal cell
  double randomness, step_cost, reward, gamma;
  int bi0, bi1, bj0, bj1;  // bounding box of the free cells (rows bi0..bi1, columns bj0..bj1; the border is wall)
};

// ---- precision-generic math ------------------------------------------------------------------
__device__ __forceinline__ float xexp(float x) { return expf(x); }
__device__ __forceinline__ double xexp(double x) { return exp(x); }
__device__ __forceinline__ float xlog(float x) { return logf(x); }
__device__ __forceinline__ double xlog(double x) { return log(x); }
__device__ __forceinline__ float xlog1p(float x) { return log1pf(x); }
__device__ __forceinline__ double xlog1p(double x) { return log1p(x); }
template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -CUDART_INF_F; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -CUDART_INF; }

// LogExpFunctions.logsumexp over the four actions (max-shifted).
template <typename T> __device__ __forceinline__ T lse4(const T (&x)[NA]) {
  T m = fmax(fmax(x[0], x[1]), fmax(x[2], x[3]));
  if (m == neg_inf<T>()) return m;
  T s = xexp(x[0] - m) + xexp(x[1] - m) + xexp(x[2] - m) + xexp(x[3] - m);
  return m + xlog(s);
}
// LogExpFunctions.logaddexp
template <typename T> __device__ __forceinline__ T logaddexp(T x, T y) {
  if (x == neg_inf<T>()) return y;
  if (y == neg_inf<T>()) return x;
  return fmax(x, y) + xlog1p(xexp(-fabs(x - y)));
}

// ---- reductions --------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// Sum over the block in a fixed order (deterministic); result valid in thread 0.
template <typename T> __device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = 0;
  if (warp == 0) {
    r = lane < nw ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// ---- mbarrier + bulk asynchronous copies (TMA engine): one thread issues, the barrier counts the bytes ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (UBLKCP): bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw; SC'11) -------------------------------------------
// The ten round keys (k0 + r*0x9E3779B9, k1 + r*0xBB67AE85) are expanded on the host and travel in
// the kernel parameters, so a round is 2 IMAD.WIDE + 2 LOP3 with constant-bank operands.
struct PhiloxKeys {
  uint32_t k[20];
};
inline PhiloxKeys philox_expand(uint32_t k0, uint32_t k1) {
  PhiloxKeys o;
  for (int r = 0; r < 10; r++) { o.k[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u; o.k[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u; }
  return o;
}
__device__ __forceinline__ uint4 philox4x32_10(const PhiloxKeys& K, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t lo0 = 0xD2511F53u * c0, hi0 = __umulhi(0xD2511F53u, c0);
    const uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = __umulhi(0xCD9E8D57u, c2);
    const uint32_t n0 = hi1 ^ c1 ^ K.k[2 * r], n2 = hi0 ^ c3 ^ K.k[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}
// 24-bit uniform in [0,1): exactly representable, identical on host and device.
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

}  // namespace piql
