// (a) cp.async.bulk 1D global->shared; (b) canonical 2D int32 tensor map example
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ void wait_bar(uint64_t* bar) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
}
__global__ void bulk1d(const unsigned char* src, unsigned char* out, int bytes) {
  __shared__ alignas(128) unsigned char sm[4096];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sm)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
  }
  wait_bar(&bar);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
__global__ void tma2d(const __grid_constant__ CUtensorMap tmap, int x, int y, int* out) {
  __shared__ alignas(128) int sm[64 * 64];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(64 * 64 * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
  }
  wait_bar(&bar);
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out[i] = sm[i];
}
int main() {
  unsigned char *d, *o;
  std::vector<unsigned char> h(1 << 20);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + 3);
  cudaMalloc(&d, h.size()); cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  bulk1d<<<1, 128>>>(d + 4096, o, 2048);
  cudaError_t e = cudaDeviceSynchronize();
  printf("bulk1d: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<unsigned char> g(2048); cudaMemcpy(g.data(), o, 2048, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 2048; ++i) bad += g[i] != h[4096 + i];
    printf("bulk1d mismatches: %d\n", bad);
  } else return 1;
  // 2D int32 1024 x 1024
  const int N = 1024;
  std::vector<int> hi((size_t)N * N); for (size_t i = 0; i < hi.size(); ++i) hi[i] = (int)i;
  int *di, *oi; cudaMalloc(&di, hi.size() * 4); cudaMalloc(&oi, 64 * 64 * 4);
  cudaMemcpy(di, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t dims[2] = {N, N};
  const cuuint64_t strides[1] = {N * 4};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, di, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode2d: %d\n", (int)r);
  unsigned char* raw = (unsigned char*)&map; printf("map bytes:"); for (int i = 0; i < 32; ++i) printf(" %02x", raw[i]); printf("\n");
  tma2d<<<1, 128>>>(map, 128, 256, oi);
  e = cudaDeviceSynchronize();
  printf("tma2d: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<int> g(64 * 64); cudaMemcpy(g.data(), oi, g.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int r2 = 0; r2 < 64; ++r2) for (int c = 0; c < 64; ++c) bad += g[r2 * 64 + c] != hi[(size_t)(256 + r2) * N + 128 + c];
    printf("tma2d mismatches: %d\n", bad);
  }
  return 0;
}
