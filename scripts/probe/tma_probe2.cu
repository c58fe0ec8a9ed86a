#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, int box_w, int box_h, unsigned char* out) {
  __shared__ alignas(128) unsigned char sm[80 * 21];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_3d_global_to_shared(sm, &tmap, x, y, z, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(sm));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = sm[i];
}
int main() {
  const int W = 10000, H = 2000, C = 4, box_w = 80, box_h = 21;
  std::vector<unsigned char> h((size_t)W * H * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
  unsigned char *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, box_w * box_h);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map;
  const cuuint64_t dims[3] = {W, H, C};
  const cuuint64_t strides[2] = {W, (cuuint64_t)W * H};
  const cuuint32_t box[3] = {box_w, box_h, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int x = 1237, y = 333, z = 2;
  probe<<<1, 128>>>(map, x, y, z, box_w, box_h, o);
  cudaError_t e = cudaDeviceSynchronize();
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
