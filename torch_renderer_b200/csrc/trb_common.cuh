// Shared helpers for the libtrb.so translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "trb.h"

namespace trb {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return TRB_ERR_CUDA;
}

#define TRB_CUDA_TRY(expr)                                 \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return ::trb::cuda_fail(_e);    \
  } while (0)

// Sets the device for the calling thread for the duration of one API call.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    want = device;
  }
  ~DeviceGuard() {
    if (ok && prev >= 0 && prev != want) cudaSetDevice(prev);
  }
  int want = -1;
};

#define TRB_ENTER(device)                                             \
  ::trb::DeviceGuard _guard(device);                                  \
  if (!_guard.ok) return ::trb::cuda_fail(cudaGetLastError());

#define TRB_LAUNCH_CHECK()                                            \
  do {                                                                \
    cudaError_t _e = cudaGetLastError();                              \
    if (_e != cudaSuccess) return ::trb::cuda_fail(_e);               \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Streaming (evict-first) stores for write-once outputs: Fragments and images are far larger
// than anything that is re-read from L2 by the same kernel.
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(long long* p, long long v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(float4* p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace trb
