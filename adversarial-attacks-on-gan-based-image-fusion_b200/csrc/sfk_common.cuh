// Shared device helpers for libsfattack (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfk.h"

#define SFK_SQRT2 1.4142135623730951f
#define SFK_RSQRT2 0.7071067811865476f

void sfk_set_error(const char* msg);
int sfk_check_launch(const char* what);

#define SFK_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) {                   \
      sfk_set_error(msg);            \
      return (code);                 \
    }                                \
  } while (0)

static inline bool sfk_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- bf16 x8 vectors (16 bytes) -------------------------------------------------------------
struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ bf16x8 ldg8(const void* p) {
  return *reinterpret_cast<const bf16x8*>(p);
}
__device__ __forceinline__ void stg8(void* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float lrelu_fwd(float v) { return (v > 0.f ? v : 0.2f * v) * SFK_SQRT2; }
// inverse of the above and its slope, both from the stored OUTPUT (sign is preserved)
__device__ __forceinline__ float lrelu_inv(float o) { return o > 0.f ? o * SFK_RSQRT2 : o * (SFK_RSQRT2 * 5.0f); }
__device__ __forceinline__ float lrelu_slope(float o) { return o > 0.f ? SFK_SQRT2 : 0.2f * SFK_SQRT2; }

static inline int sfk_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
