// peer.cu -- the multi-GPU exchange step of the Monte-Carlo estimator as ONE kernel over NVLink peer memory.
//
// Trajectories are sharded over the GPUs of a box by global path id; every rank ends an iteration with partial
// log-sum-exp statistics (max, sum exp(x - max), count) for every (start slot, first action).  The reference has
// no such step (its Distributed.jl pool -- addprocs/pmap/@spawn, scripts/distributed_experiments.jl:3,26,28,43-46
// -- is task parallel); with NCCL it takes three collectives around a rescale kernel because log-sum-exp is not a
// reduction operator NCCL knows.  Here each rank owns the slice [rank*n/W, (rank+1)*n/W) of the slots: it pulls
// that slice of all three arrays from every peer with 128-bit loads over NVLink, combines the W triples exactly
// (global max first, then one rescaled sum in rank order), and pushes the result into every rank's arrays --
// reduce-scatter + all-gather in one pass, no temporary, no second kernel, identical bits on every rank.
// Link traffic per rank and direction: (W-1)/W of the 12 B x n accumulators twice (read responses for the pulled
// slices + the peers' pushes); measured on 2 x B200: 642 MB in 1.05 ms = 612 GB/s per direction (peer-copy
// reference 770 GB/s); pull alone 0.60 ms, push alone 0.56 ms -- they share each direction, so they add up.
//
// Cross-GPU ordering uses two flag words per (rank, peer) in peer-visible memory, written with system-scope
// fences: an entry barrier (everyone's local reduction has finished -- stream order on each device) and an exit
// barrier held by the last block to finish (every slice has landed everywhere; nobody still reads my inputs).
// Each rank's kernel is alone on its own GPU, so the kernels do run concurrently; waits are bounded by a
// wall-clock limit and report through an error word instead of hanging the device.
#include <cstring>

#include "internal.h"

namespace piql {

struct PeerArgs {
  float* mx[PEER_MAX];
  float* sm[PEER_MAX];
  uint32_t* cn[PEER_MAX];
  unsigned long long* flags[PEER_MAX];  // flags[r] lives on rank r: [phase][source rank]
  int rank, world;
  long long nslots;
  unsigned long long epoch;
  unsigned int* done;  // local: blocks that finished their part
  int* err;            // local: set when a wait timed out
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// thread r < world tells rank r "I reached `phase` of `epoch`"
__device__ __forceinline__ void peer_signal(const PeerArgs& a, int phase) {
  __threadfence_system();
  if ((int)threadIdx.x < a.world) {
    volatile unsigned long long* f = a.flags[threadIdx.x] + phase * PEER_MAX + a.rank;
    *f = a.epoch;
  }
}
// thread r < world waits until rank r has signalled; bounded (10 s)
__device__ __forceinline__ void peer_wait(const PeerArgs& a, int phase) {
  if ((int)threadIdx.x < a.world) {
    volatile unsigned long long* f = a.flags[a.rank] + phase * PEER_MAX + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (*f < a.epoch) {
      if (globaltimer_ns() - t0 > 10000000000ull) { *a.err = 1 + phase; break; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

__device__ __forceinline__ float4 ldcv4(const float* p) { return __ldcv(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint4 ldcv4(const uint32_t* p) { return __ldcv(reinterpret_cast<const uint4*>(p)); }

// exp(m - M) with the empty-bucket convention (max = -inf, sum = 0)
__device__ __forceinline__ float scale_to(float m, float M) { return m == -CUDART_INF_F ? 0.f : expf(m - M); }

template <int W>
__global__ void __launch_bounds__(256)
lse_allreduce_peer_kernel(const PeerArgs a) {
  __shared__ bool last;
  if (blockIdx.x == 0) peer_signal(a, 0);  // my accumulators are final (stream order put my reduction before me)
  peer_wait(a, 0);
  __syncthreads();

  const long long s0 = a.nslots * a.rank / a.world, s1 = a.nslots * (a.rank + 1) / a.world;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; s < s1; s += stride) {
    float4 m[W], v[W];
    uint4 c[W];
#pragma unroll
    for (int r = 0; r < W; r++) {  // all loads in flight before the first use
      m[r] = ldcv4(a.mx[r] + 4 * s);
      v[r] = ldcv4(a.sm[r] + 4 * s);
      c[r] = ldcv4(a.cn[r] + 4 * s);
    }
    float4 M = m[0];
    uint4 C = c[0];
#pragma unroll
    for (int r = 1; r < W; r++) {
      M.x = fmaxf(M.x, m[r].x); M.y = fmaxf(M.y, m[r].y); M.z = fmaxf(M.z, m[r].z); M.w = fmaxf(M.w, m[r].w);
      C.x += c[r].x; C.y += c[r].y; C.z += c[r].z; C.w += c[r].w;
    }
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < W; r++) {  // fixed rank order: every rank would compute the same bits
      S.x += v[r].x * scale_to(m[r].x, M.x); S.y += v[r].y * scale_to(m[r].y, M.y);
      S.z += v[r].z * scale_to(m[r].z, M.z); S.w += v[r].w * scale_to(m[r].w, M.w);
    }
#pragma unroll
    for (int r = 0; r < W; r++) {
      *reinterpret_cast<float4*>(a.mx[r] + 4 * s) = M;
      *reinterpret_cast<float4*>(a.sm[r] + 4 * s) = S;
      *reinterpret_cast<uint4*>(a.cn[r] + 4 * s) = C;
    }
  }

  // exit barrier: the last block of this rank tells the peers that its slice has landed, then holds the kernel
  // until every peer has said the same (so finalize sees all slices, and a later reset cannot race a peer's read)
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *a.done = 0u;
    peer_signal(a, 1);
    peer_wait(a, 1);
  }
}

// ---- host ----------------------------------------------------------------------------------------------------
void launch_lse_allreduce_peer(const PeerState& ps, long long nslots, int num_sms, cudaStream_t st) {
  PeerArgs a{};
  for (int r = 0; r < ps.world; r++) {
    a.mx[r] = ps.mx[r]; a.sm[r] = ps.sm[r]; a.cn[r] = ps.cn[r]; a.flags[r] = ps.flags[r];
  }
  a.rank = ps.rank; a.world = ps.world; a.nslots = nslots; a.epoch = ps.epoch;
  a.done = ps.done; a.err = ps.err;
  const long long slice = nslots / ps.world + 1;
  long long want = (slice + 255) / 256;
  const long long cap = 4LL * num_sms;  // measured: 2x..16x SMs all give the same time; the links are the limit
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  switch (ps.world) {
    case 1: lse_allreduce_peer_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 2: lse_allreduce_peer_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 3: lse_allreduce_peer_kernel<3><<<grid, 256, 0, st>>>(a); break;
    case 4: lse_allreduce_peer_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 5: lse_allreduce_peer_kernel<5><<<grid, 256, 0, st>>>(a); break;
    case 6: lse_allreduce_peer_kernel<6><<<grid, 256, 0, st>>>(a); break;
    case 7: lse_allreduce_peer_kernel<7><<<grid, 256, 0, st>>>(a); break;
    default: lse_allreduce_peer_kernel<8><<<grid, 256, 0, st>>>(a); break;
  }
  g_launches++;
}

}  // namespace piql
