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

// ---- bf16 x8 vectors (16 bytes) --------------------------------------------------------------
// Carried as uint4 so that every access is ONE 128-bit LDG/STG (a struct of four bfloat162 gets scalarised
// into four 32-bit accesses by nvcc); bf16 <-> fp32 is a 16-bit shift.
typedef uint4 bf16x8;

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  f[0] = __uint_as_float(p.x << 16);
  f[1] = __uint_as_float(p.x & 0xffff0000u);
  f[2] = __uint_as_float(p.y << 16);
  f[3] = __uint_as_float(p.y & 0xffff0000u);
  f[4] = __uint_as_float(p.z << 16);
  f[5] = __uint_as_float(p.z & 0xffff0000u);
  f[6] = __uint_as_float(p.w << 16);
  f[7] = __uint_as_float(p.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
  p.x = pack2(f[0], f[1]);
  p.y = pack2(f[2], f[3]);
  p.z = pack2(f[4], f[5]);
  p.w = pack2(f[6], f[7]);
  return p;
}
// read-only (non-coherent) 128-bit load: only for buffers this kernel never writes
__device__ __forceinline__ bf16x8 ldg8(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// plain 128-bit load: for buffers updated in place by the same kernel
__device__ __forceinline__ bf16x8 ld8(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg8(void* p, const bf16x8& v) { *reinterpret_cast<uint4*>(p) = v; }

// ---- storage-type generic 8-element vectors: activations are bf16 (product path) or fp32 (parity mode, sfk_set_activation_dtype)
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) { unpack8(ldg8(p), f); }
__device__ __forceinline__ void load8p(const __nv_bfloat16* p, float* f) { unpack8(ld8(p), f); }
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) { stg8(p, pack8(f)); }
__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load8p(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ void from_f32(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void from_f32(float* p, float v) { *p = v; }

// Streaming (cp.async.bulk ring) form of sfk_act_bwd (wrgb == nullptr) / sfk_act_torgb_bwd, csrc/sfk_stream.cu.
// Returns -1000 when the shape is outside its domain (the caller then launches the register kernel), else the launch status.
int sfk_act_stream_launch(const void* out, const void* gin, void* gz, const float* d, const float* noise, float noise_w, const float* bias,
                          float* gdacc, const float* wrgb, const float* s_rgb, int s_stride, const float* grgb, float* gs_rgb, int gs_stride,
                          const float* s_in, float* gs_in, int in_stride, int gin_stride, int n, int hw, int c, cudaStream_t st);

int sfk_torgb_stream_launch(const void* x, const float* wrgb, const float* s, int s_stride, const float* bias, const float* skip, float* rgb,
                            int n, int h, int w, int c, cudaStream_t st);

// Streaming form of sfk_blur_act_fwd / sfk_blur_act_bwd, csrc/sfk_blur_stream.cu (same -1000 convention)
int sfk_blur_stream_launch(bool bwd, const void* a0, const void* a1, void* dst, const float* d, const float* noise, float noise_w, const float* bias,
                           float* gdacc, const float* s_in, float* gs_in, int in_stride, int n, int h, int w, int c, cudaStream_t st);

// activation storage mode of the library: 0 = bf16 (default), 1 = fp32 (parity mode; tensor-core conv unavailable)
int sfk_act_f32();
#define SFK_ACT_DISPATCH(CALL_BF16, CALL_F32) \
  do {                                          \
    if (sfk_act_f32()) {                        \
      CALL_F32;                                 \
    } else {                                    \
      CALL_BF16;                                \
    }                                           \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float lrelu_fwd(float v) { return (v > 0.f ? v : 0.2f * v) * SFK_SQRT2; }
// inverse of the above and its slope, both from the stored OUTPUT (sign is preserved)
__device__ __forceinline__ float lrelu_inv(float o) { return o > 0.f ? o * SFK_RSQRT2 : o * (SFK_RSQRT2 * 5.0f); }
__device__ __forceinline__ float lrelu_slope(float o) { return o > 0.f ? SFK_SQRT2 : 0.2f * SFK_SQRT2; }

// SM count of the CURRENT device (cached per device ordinal: a process may drive several GPUs)
static inline int sfk_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = sms[dev & 63];
  if (!v) {
    int q = 0;
    cudaDeviceGetAttribute(&q, cudaDevAttrMultiProcessorCount, dev);
    v = q > 0 ? q : 148;
  }
  return v;
}
