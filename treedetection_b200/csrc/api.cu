// libtreedet C-ABI plumbing: version, thread-local error string.
// Every entry point is declared in include/treedet.h.
#include <cstdarg>

#include "common.cuh"

static thread_local char g_err[512] = "";

void td_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Temporary scratch comes from the device's default stream-ordered pool.  Its default release
// threshold is 0, i.e. every stream synchronisation hands the freed blocks back to the OS and
// the next call pays for mapping them again (milliseconds); keep them cached instead.
void td_ensure_pool() {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}

static thread_local TdArena* g_arena = nullptr;

void td_set_arena(TdArena* arena) {
  g_arena = arena;
  if (arena) arena->off = 0;
}

void td_arena_rewind() {
  if (g_arena) g_arena->off = 0;
}

cudaError_t td_tmp_alloc(void** p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) bytes = 1;
  if (g_arena) {
    const size_t o = (g_arena->off + 255) & ~(size_t)255;
    if (o + bytes <= g_arena->bytes) {
      *p = g_arena->base + o;
      g_arena->off = o + bytes;
      return cudaSuccess;
    }
  }
  td_ensure_pool();
  return cudaMallocAsync(p, bytes, st);
}

void td_tmp_free(void* p, cudaStream_t st) {
  if (!p) return;
  if (g_arena && (char*)p >= g_arena->base && (char*)p < g_arena->base + g_arena->bytes) return;
  cudaFreeAsync(p, st);
}

extern "C" int td_version() { return 100; }  // 0.1.0

extern "C" const char* td_last_error() { return g_err; }

// number of SMs of the current device (148 on B200); also proves the CUDA runtime
// is usable from the calling process
extern "C" int td_device_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { td_set_error("cudaGetDevice failed"); return TD_ERR_CUDA; }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    td_set_error("cudaDeviceGetAttribute failed");
    return TD_ERR_CUDA;
  }
  return n;
}
