// C-ABI plumbing: error string, version, device queries.
#include <stdarg.h>
#include <string.h>

#include "host.h"

namespace emap {

static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

// A zeroed 32-bit counter in device memory for one kernel launch on `stream` (dynamic tile scheduling of the
// persistent MLP kernels): a per-device ring of 64 counters, zeroed in stream order right before the launch.
// Returns NULL when no counter can be had (the kernels then fall back to the static round robin).
unsigned int* tile_counter(cudaStream_t stream) {
  static unsigned int* pool[64] = {nullptr};
  static unsigned int next[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (!pool[dev]) {
    if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return nullptr;
    if (cudaMalloc(&pool[dev], 64 * 64) != cudaSuccess) { pool[dev] = nullptr; cudaGetLastError(); return nullptr; }
  }
  unsigned int* c = pool[dev] + 16 * (next[dev]++ & 63);       // 64 bytes apart
  if (cudaMemsetAsync(c, 0, sizeof(unsigned int), stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return c;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace emap

extern "C" const char* emap_last_error(void) { return emap::g_err; }
extern "C" int emap_abi_version(void) { return EMAP_ABI_VERSION; }
