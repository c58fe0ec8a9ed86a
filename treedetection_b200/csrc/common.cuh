// Shared helpers for libtreedet (sm_100a).  No torch types anywhere: the
// library sees raw device pointers, sizes and a cudaStream_t.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define TD_HD __host__ __device__
#define TD_D __device__ __forceinline__
#else
// Host-simulation build (tests/hostsim): the per-item sequential algorithms
// (contour tracing, ring simplification, exact predicates) are plain C++ and
// can be compiled by g++ so that they are debuggable without a GPU.  The
// product never uses that build.
#define TD_HD
#endif

// ---- error codes (include/treedet.h) --------------------------------------
#define TD_OK 0
#define TD_ERR_CUDA (-1)
#define TD_ERR_ARG (-2)
#define TD_ERR_OVERFLOW (-3)
#define TD_ERR_UNSUPPORTED (-4)

#if defined(__CUDACC__)
void td_set_error(const char* fmt, ...);
void td_ensure_pool();   // keeps the stream-ordered scratch pool cached across synchronisations

// Temporary device scratch of one entry point.  By default it comes from the stream-ordered pool
// (cudaMallocAsync / cudaFreeAsync).  td_chain_* installs an ARENA for the calling thread instead: a region of
// its workspace that every composed call bumps through from the start (the calls are ordered on one
// stream, so a region is dead by the time the next call's kernels run).  That keeps memory-allocation
// nodes out of the captured CUDA graphs -- a graph that owns allocations re-maps them on launch, which
// showed up as multi-millisecond stalls in front of its first memset node.
struct TdArena {
  char* base;
  size_t bytes;
  size_t off;
};
void td_set_arena(TdArena* arena);            // nullptr: back to the stream-ordered pool
void td_arena_rewind();                       // the next td_tmp_alloc starts at the arena's beginning again
cudaError_t td_tmp_alloc(void** p, size_t bytes, cudaStream_t st);
void td_tmp_free(void* p, cudaStream_t st);

#define TD_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      td_set_error("%s: %s", name, cudaGetErrorString(e__));                         \
      return TD_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define TD_CUDA(call)                                                                \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      td_set_error("%s: %s", #call, cudaGetErrorString(e__));                        \
      return TD_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define TD_ARG(cond)                                                                 \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      td_set_error("%s: bad argument: %s", __func__, #cond);                         \
      return TD_ERR_ARG;                                                             \
    }                                                                                \
  } while (0)

static inline int td_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// B200: 148 SMs.  Grid-stride kernels are launched with a multiple of this.
static inline int td_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
#endif
