// Minimal TMA 2-D box load probe (debug tool): nvcc -gencode arch=compute_100a,code=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename T>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int EJ, T* out, int variant) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long mbar;
  T* buf = (T*)sm;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    if (variant & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"((unsigned)(EJ * 128 * sizeof(T))) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(buf)), "l"(&tm), "r"(c0), "r"(c1), "r"(smem_u32(&mbar)) : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&mbar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < EJ * 128; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <typename T> int run(CUtensorMapDataType dt, int EJ, int variant) {
  const int ld = 1088, Wp = 200;
  std::vector<T> h((size_t)ld * Wp);
  for (size_t i = 0; i < h.size(); i++) h[i] = (T)(i % 1000003);
  T *d, *o;
  cudaMalloc(&d, sizeof(T) * h.size());
  cudaMalloc(&o, sizeof(T) * 128 * EJ);
  cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)Wp}; cuuint64_t gs[1] = {(cuuint64_t)ld * sizeof(T)};
  cuuint32_t box[2] = {128, (cuuint32_t)EJ}; cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&tm, dt, 2, d, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("sizeof(T)=%zu EJ=%d variant=%d encode=%d ", sizeof(T), EJ, variant, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  cudaFuncSetAttribute(probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<T><<<1, 256, EJ * 128 * sizeof(T)>>>(tm, 120, 40, EJ, o, variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run=%s ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 2; }
  std::vector<T> got(128 * EJ);
  cudaMemcpy(got.data(), o, sizeof(T) * got.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < EJ; j++) for (int i = 0; i < 128; i++) bad += got[j * 128 + i] != h[(size_t)(40 + j) * ld + 120 + i];
  printf("mismatches=%d\n", bad);
  return 0;
}
int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) return run<float>(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 16, 1);
  if (which == 1) return run<float>(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 90, 1);
  if (which == 2) return run<double>(CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 16, 1);
  if (which == 3) return run<uint8_t>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 16, 1);
  if (which == 4) return run<float>(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 16, 0);
  if (which == 5) return run<float>(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 64, 1);
  return 0;
}
