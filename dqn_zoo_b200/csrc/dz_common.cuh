// Shared host/device helpers for the dqn_zoo_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/dqn_zoo_b200.h"

namespace dz {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  char buf[512];
  snprintf(buf, sizeof(buf), fmt, a, b);
  g_last_error = buf;
  return code;
}

#define DZ_CUDA_OK(expr)                                                                     \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return dz::fail(DZ_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// Optional per-launch CUDA-event timing (dz_profile_begin/end): used by bench.py to time the
// dominant kernel on its own stream.  Off in every timed run.
extern bool g_profile;
void profile_mark(const char* name, void* stream, bool begin);

// Every kernel launch goes through this so bench.py can report `gpu_launches`.
#define DZ_LAUNCH_NAMED(name, kernel, grid, block, smem, stream, ...)                        \
  do {                                                                                       \
    if (dz::g_profile) dz::profile_mark(name, stream, true);                                 \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                \
    if (dz::g_profile) dz::profile_mark(name, stream, false);                                \
    dz::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) return dz::fail(DZ_ECUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)
#define DZ_LAUNCH(kernel, grid, block, smem, stream, ...) \
  DZ_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace dz
