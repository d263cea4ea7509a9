// Shared helpers for the fedcola_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fedcola_b200.h"   // FC_OK / FC_ERR_* and the entry-point declarations

// Last error text, per thread (entry points are called from the reference's ThreadPoolExecutor workers).
extern thread_local char fc_last_error_buf[512];

#define FC_FAIL(code, ...)                                              \
  do {                                                                  \
    snprintf(fc_last_error_buf, sizeof(fc_last_error_buf), __VA_ARGS__); \
    return (code);                                                      \
  } while (0)

#define FC_CUDA_CHECK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      FC_FAIL(FC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Every kernel launch of the library goes through FC_LAUNCH_CHECK: it also feeds fc_launch_count().
extern unsigned long long fc_launch_counter;
#define FC_LAUNCH_CHECK()                                         \
  do {                                                            \
    __atomic_fetch_add(&fc_launch_counter, 1ULL, __ATOMIC_RELAXED); \
    FC_CUDA_CHECK(cudaGetLastError());                            \
  } while (0)

#define FC_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) FC_FAIL(FC_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Test / measurement hook (fc_set_grid_cap): upper bound on the CTA count of the persistent kernels, so that parity
// tests can force many tiles / items per CTA on small problems.  0 = no cap.
extern int fc_grid_cap_value;
static inline int fc_apply_grid_cap(int grid) {
  const int cap = __atomic_load_n(&fc_grid_cap_value, __ATOMIC_RELAXED);
  return (cap > 0 && grid > cap) ? cap : grid;
}

static inline int fc_num_sms(int device) {
  static thread_local int cached_dev = -1, cached = 0;
  if (cached_dev != device) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, device);
    cached_dev = device;
  }
  return cached > 0 ? cached : 148;
}

// Opt a kernel into > 48 KB of dynamic shared memory once per (thread, device) instead of on every launch.
#define FC_SMEM_OPT_IN(kernel, bytes)                                                             \
  do {                                                                                            \
    static thread_local int _cfg_dev = -1;                                                        \
    static thread_local int _cfg_bytes = 0;                                                       \
    int _dev = -1;                                                                                \
    cudaGetDevice(&_dev);                                                                         \
    if (_dev != _cfg_dev || (int)(bytes) > _cfg_bytes) {                                          \
      FC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _cfg_dev = _dev;                                                                            \
      _cfg_bytes = (int)(bytes);                                                                  \
    }                                                                                             \
  } while (0)

struct FcDeviceGuard {
  int prev;
  explicit FcDeviceGuard(int dev) { cudaGetDevice(&prev); if (dev >= 0 && dev != prev) cudaSetDevice(dev); else prev = -1; }
  ~FcDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// 128-bit streaming loads/stores that bypass L1 (data touched once).
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

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
