// standalone probe: 3D TMA box load of a (W,H,C) uint8 tensor, copy back, compare on host
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, int box_w, int box_h, unsigned char* out, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    if (mode & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((uint32_t)(box_w * box_h)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(sm)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  }
  for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 1; int bw = argc > 2 ? atoi(argv[2]) : 80; int bh = argc > 3 ? atoi(argv[3]) : 21;
  const int W = 10000, H = 2000, C = 4; const int box_w = bw, box_h = bh;
  std::vector<unsigned char> h((size_t)W * H * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
  unsigned char *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, box_w * box_h);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, p);
  CUtensorMap map;
  const cuuint64_t dims[3] = {W, H, C};
  const cuuint64_t strides[2] = {W, (cuuint64_t)W * H};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, (mode & 2) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int x = 1237, y = 333, z = 2;
  probe<<<1, 256, 8192 + 64>>>(map, x, y, z, box_w, box_h, o, mode);
  e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<unsigned char> g(box_w * box_h);
  cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < box_h; ++r2) for (int c = 0; c < box_w; ++c)
    if (g[r2 * box_w + c] != h[(size_t)z * W * H + (size_t)(y + r2) * W + x + c]) ++bad;
  printf("mismatches: %d\n", bad);
  return 0;
}
