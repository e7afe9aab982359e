// Shared helpers for libb200gen.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200gen.h"

namespace b200 {

// thread-local last-error text, set by every failing entry point
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B200_CHECK_ARG(cond, ...)                    \
  do {                                               \
    if (!(cond)) {                                   \
      b200::set_error(__VA_ARGS__);                  \
      return B200_EINVAL;                            \
    }                                                \
  } while (0)

#define B200_CUDA(call)                                              \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return b200::cuda_fail(e__, #call);      \
  } while (0)

#define B200_LAUNCH_CHECK(name)                                      \
  do {                                                               \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) return b200::cuda_fail(e__, name);       \
  } while (0)

int sm_count();

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == B200_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == B200_ACT_SILU) return silu_f(x);
  if (act == B200_ACT_LEAKYRELU) return x > 0.0f ? x : 0.01f * x;
  if (act == B200_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}

// 8 bf16 <-> 8 floats through one 16-byte vector
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
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

}  // namespace b200
