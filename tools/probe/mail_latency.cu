// mail_latency.cu -- one-way latency of an L2 "mailbox" between two CTAs on different SMs (sm_100a), for the access flavours the
// column-resident sweep kernel could use.  Ping-pong: CTA 0 posts seq to box A, CTA 1 polls A and posts to box B, CTA 0 polls B.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mail_latency mail_latency.cu && ./mail_latency
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int MODE> __device__ __forceinline__ void post(unsigned long long* p, unsigned long long a, unsigned long long b) {
  if (MODE == 0) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
  if (MODE == 1) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
  if (MODE == 2) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
  if (MODE == 3) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory"); asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p + 1), "l"(b) : "memory"); }
}
template <int MODE> __device__ __forceinline__ void peek(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  if (MODE == 0) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if (MODE == 3) { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory"); asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(b) : "l"(p + 1) : "memory"); }
}

// pairs of CTAs (2i, 2i+1) ping-pong; lanes = threads of the CTA that post / poll their own 16 bytes
template <int MODE> __global__ void pingpong(unsigned long long* box, int iters, int stride_ctas, long long* cycles) {
  const int pair = blockIdx.x % stride_ctas, side = blockIdx.x / stride_ctas;  // partner = blockIdx.x +- stride_ctas (another SM)
  unsigned long long* A = box + ((size_t)pair * 2 + 0) * 2 * blockDim.x + 2 * threadIdx.x;
  unsigned long long* B = box + ((size_t)pair * 2 + 1) * 2 * blockDim.x + 2 * threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 1; i <= iters; i++) {
    const unsigned long long v = ((unsigned long long)i << 32) | threadIdx.x;
    unsigned long long a, b;
    if (side == 0) {
      post<MODE>(A, v, v);
      do { peek<MODE>(B, a, b); } while ((unsigned)(a >> 32) != (unsigned)i || (unsigned)(b >> 32) != (unsigned)i);
    } else {
      do { peek<MODE>(A, a, b); } while ((unsigned)(a >> 32) != (unsigned)i || (unsigned)(b >> 32) != (unsigned)i);
      post<MODE>(B, v, v);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the same between the two CTAs of a cluster through distributed shared memory
__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int iters, long long* cycles) {
  __shared__ unsigned long long boxs[2 * 1024];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  boxs[2 * threadIdx.x] = 0; boxs[2 * threadIdx.x + 1] = 0;
  cl.sync();
  volatile unsigned long long* mine = boxs + 2 * threadIdx.x;
  unsigned long long* theirs = cl.map_shared_rank(boxs, rank ^ 1) + 2 * threadIdx.x;
  const long long t0 = clock64();
  for (int i = 1; i <= iters; i++) {
    const unsigned long long v = ((unsigned long long)i << 32) | threadIdx.x;
    if (rank == 0) {
      *(volatile unsigned long long*)theirs = v;
      while ((unsigned)(*mine >> 32) != (unsigned)i) {}
    } else {
      while ((unsigned)(*mine >> 32) != (unsigned)i) {}
      *(volatile unsigned long long*)theirs = v;
    }
  }
  const long long t1 = clock64();
  cl.sync();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int pairs, int threads, unsigned long long* box, long long* dcyc, int clock_khz) {
  const int iters = 2000;
  cudaMemset(box, 0, (size_t)pairs * 2 * 2 * threads * 8);
  void* args[] = {&box, (void*)&iters, (void*)&pairs, &dcyc};
  cudaLaunchCooperativeKernel((void*)pingpong<MODE>, dim3(2 * pairs), dim3(threads), args, 0, 0);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[512];
  cudaMemcpy(h, dcyc, sizeof(long long) * 2 * pairs, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < 2 * pairs; i++) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("%-28s pairs %3d threads %4d: one-way %7.0f .. %7.0f cycles (%.2f us at %d MHz) %s\n", name, pairs, threads,
         mn / (2.0 * iters), mx / (2.0 * iters), mx / (2.0 * iters) / (clock_khz * 1e-3), clock_khz / 1000, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  unsigned long long* box;
  long long* dcyc;
  cudaMalloc(&box, 64 << 20);
  cudaMalloc(&dcyc, 8 * 512);
  printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
  const int half = p.multiProcessorCount / 2;
  for (int threads : {32, 512}) {
    for (int pairs : {1, half}) {
      run<0>("volatile v2.u64", pairs, threads, box, dcyc, khz);
      run<1>("relaxed.gpu v2.u64", pairs, threads, box, dcyc, khz);
      run<2>("cg v2.u64", pairs, threads, box, dcyc, khz);
      run<3>("relaxed.gpu 2 x u64", pairs, threads, box, dcyc, khz);
    }
  }
  for (int threads : {32, 512}) {
    const int iters = 2000;
    pingpong_dsmem<<<2, threads>>>(iters, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, dcyc, 16, cudaMemcpyDeviceToHost);
    printf("dsmem cluster pair threads %4d: one-way %7.0f cycles %s\n", threads, (h[0] > h[1] ? h[0] : h[1]) / (2.0 * iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
