// generic probe: dtype (1 or 4 bytes), rank 2/3, dims, box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int rank, int x, int y, int z, int bytes, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (rank == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(sm)), "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(sm)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  const int es = atoi(argv[1]), rank = atoi(argv[2]);
  const long W = atol(argv[3]), H = atol(argv[4]), C = atol(argv[5]);
  const int bw = atoi(argv[6]), bh = atoi(argv[7]);
  const int x = atoi(argv[8]), y = atoi(argv[9]), z = atoi(argv[10]);
  std::vector<unsigned char> h((size_t)W * H * C * es);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
  unsigned char *d, *o;
  const int bytes = bw * bh * es;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, bytes);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C};
  const cuuint64_t strides[2] = {(cuuint64_t)W * es, (cuuint64_t)W * H * es};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t est[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)p)(&map, es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, rank, d, dims, strides, box, est,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("es=%d rank=%d W=%ld H=%ld C=%ld box=%dx%d at (%d,%d,%d): encode %d ", es, rank, W, H, C, bw, bh, x, y, z, (int)r);
  k<<<1, 128, bytes + 128>>>(map, rank, x, y, z, bytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 1; }
  std::vector<unsigned char> g(bytes);
  cudaMemcpy(g.data(), o, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2) for (int c = 0; c < bw * es; ++c)
    if (g[r2 * bw * es + c] != h[((size_t)z * W * H + (size_t)(y + r2) * W + x) * es + c]) ++bad;
  printf("mismatches: %d\n", bad);
  return 0;
}
