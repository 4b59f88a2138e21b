// Streaming form of the bandwidth-bound activation-backward kernels for the high-resolution synthesis layers.
//
// The register-only kernels (sfk_elementwise.cu: act_bwd_kernel, act_torgb_bwd_kernel) keep one or two 16-byte loads per
// thread in flight: at ~100 registers per thread that is ~32 KB per SM, which by Little's law (HBM latency x 6.5 TB/s / 148 SMs
// ~ 35 KB per SM) cannot fill the memory pipe -- measured 2.5 TB/s on the 32-channel 1024^2 layer.  Here the loads are decoupled
// from the registers: a producer warp streams 8 KB chunks of every input tensor into a 4-stage shared-memory ring with
// cp.async.bulk (1-D TMA, completion on an mbarrier), eight consumer warps read the ring with conflict-free LDS.128, do the
// arithmetic and store the result straight to global memory (16 B per thread, 512 contiguous bytes per warp).  Two CTAs per SM
// keep 2 x 4 x 16 KB = 128 KB of loads in flight per SM.
//
// Arithmetic is identical to the register kernels (same per-element expression order), so both forms are interchangeable;
// the planner (act_stream_launch) picks this one for bf16 storage when a layer is large enough to amortise the ring start-up.
#include <stdlib.h>

#include "sfk_common.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kChunkBytes = 8192;                 // per streamed activation tensor and stage
constexpr int kConsumers = 256;                   // 8 consumer warps
constexpr int kThreads = kConsumers + 32;         // + 1 producer warp
constexpr int kItems = kChunkBytes / 16 / kConsumers;   // 16-byte vectors per consumer thread and chunk (2)
constexpr int kAuxBytes = 2048;                   // per stage: 3 planes of grgb + noise, <= 128 pixels x 4 B each
constexpr int kStageBytes = 2 * kChunkBytes + kAuxBytes;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mb_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(s_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded: a ring that never advances traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
  for (long i = 0; i < (1L << 28); ++i)
    if (mb_try(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

struct ActStreamK {
  const __nv_bfloat16* out;
  const __nv_bfloat16* gin;   // may alias gz
  __nv_bfloat16* gz;
  const float* d;
  const float* noise;         // [HW] or null
  float noise_w;
  const float* bias;
  float* gdacc;
  const float* wrgb;          // [3][C]         (RGB)
  const float* s_rgb;         // + n * s_stride (RGB)
  const float* grgb;          // [n][3][HW]     (RGB)
  float* gs_rgb;              // + n * gs_stride (RGB)
  const float* s_in;          // + n * in_stride, or null
  float* gs_in;               // + n * gin_stride, or null
  int s_stride, gs_stride, in_stride, gin_stride;
  int HW, C, P, chunks, vshift;   // P pixels per chunk, chunks per image, vshift = log2(C/8)
};

// RGB: the ToRGB backward rides along (act_torgb_bwd); GIN: an incoming gradient exists (false only at the top resolution)
template <bool RGB, bool GIN>
__global__ void __launch_bounds__(kThreads, 2) act_stream_kernel(const __grid_constant__ ActStreamK a) {
  extern __shared__ __align__(128) uint8_t sm_raw[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const uint32_t ring = (s_u32(sm_raw) + 127u) & ~127u;
  float* const scratch = reinterpret_cast<float*>(sm_raw + (ring - s_u32(sm_raw)) + kStages * kStageBytes);   // [C] accumulators, then [C/8][3][8] weights
  float* const sacc = scratch;
  float* const sw = scratch + a.C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y;
  const int C = a.C, P = a.P;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mb_init(&full_bar[i], 1);
      mb_init(&empty_bar[i], kConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (RGB) {
    for (int i = tid; i < 3 * C; i += kThreads) {
      const int col = i / C, c = i % C;
      sw[((c >> 3) * 3 + col) * 8 + (c & 7)] = a.wrgb[i];
    }
  }
  __syncthreads();
  const long img = static_cast<long>(n) * a.HW;

  float racc[8], rrgb[8], rin[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) racc[i] = rrgb[i] = rin[i] = 0.f;
  const int vecs = C >> 3;
  const int cv = tid & (vecs - 1);

  if (warp == kConsumers / 32) {
    // ===================== producer: one lane issues the bulk copies of a chunk =====================
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(kChunkBytes * (GIN ? 2 : 1) + (RGB ? 3 : 0) * P * 4 + (a.noise ? P * 4 : 0));
      int it = 0;
      for (int ch = blockIdx.x; ch < a.chunks; ch += gridDim.x, ++it) {
        const int st = it % kStages;
        mb_wait(&empty_bar[st], ((it / kStages) & 1) ^ 1);
        mb_expect_tx(&full_bar[st], tx);
        const long p0 = static_cast<long>(ch) * P;
        const uint32_t base = ring + st * kStageBytes;
        bulk_g2s(base, a.out + (img + p0) * C, kChunkBytes, &full_bar[st]);
        if (GIN) bulk_g2s(base + kChunkBytes, a.gin + (img + p0) * C, kChunkBytes, &full_bar[st]);
        if (RGB) {
          const float* g0 = a.grgb + static_cast<long>(n) * 3 * a.HW + p0;
#pragma unroll
          for (int c = 0; c < 3; ++c) bulk_g2s(base + 2 * kChunkBytes + c * P * 4, g0 + static_cast<long>(c) * a.HW, P * 4, &full_bar[st]);
        }
        if (a.noise) bulk_g2s(base + 2 * kChunkBytes + 3 * P * 4, a.noise + p0, P * 4, &full_bar[st]);
      }
    }
  } else {
    // ===================== consumers =====================
    float sv[8], bv[8], si[8], kp[8];
    const float4* const wq = reinterpret_cast<const float4*>(sw + cv * 24);   // ToRGB weights of this thread's 8 channels: [3][8]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cv * 8 + i;
      sv[i] = RGB ? a.s_rgb[static_cast<long>(n) * a.s_stride + c] : 0.f;
      bv[i] = a.bias[c];
      si[i] = a.s_in ? a.s_in[static_cast<long>(n) * a.in_stride + c] : 1.f;
      kp[i] = a.d[static_cast<long>(n) * C + c] * SFK_SQRT2;      // gz = g * kp * (out > 0 ? 1 : 0.2)
    }
    int it = 0;
    for (int ch = blockIdx.x; ch < a.chunks; ch += gridDim.x, ++it) {
      const int st = it % kStages;
      const uint32_t base = ring + st * kStageBytes;
      const long p0 = static_cast<long>(ch) * P;
      mb_wait(&full_bar[st], (it / kStages) & 1);
      uint4 ovr[kItems], gvr[kItems];
      float ga[kItems], gb[kItems], gc[kItems], nz[kItems];
#pragma unroll
      for (int k = 0; k < kItems; ++k) {      // every shared-memory read of the chunk first, so the slot can be released early
        const int item = tid + k * kConsumers;
        const int p = item >> a.vshift;
        ovr[k] = lds128(base + item * 16);
        gvr[k] = GIN ? lds128(base + kChunkBytes + item * 16) : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t aux = base + 2 * kChunkBytes + p * 4;
        ga[k] = RGB ? lds32(aux) : 0.f;
        gb[k] = RGB ? lds32(aux + P * 4) : 0.f;
        gc[k] = RGB ? lds32(aux + 2 * P * 4) : 0.f;
        nz[k] = a.noise ? a.noise_w * lds32(aux + 3 * P * 4) : 0.f;
      }
      __syncwarp();
      if (lane == 0) mb_arrive(&empty_bar[st]);
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        const int item = tid + k * kConsumers;
        const int p = item >> a.vshift;
        float ov[8], gv[8], o[8];
        unpack8(ovr[k], ov);
        unpack8(gvr[k], gv);
        float w0[8], w1[8], w2[8];
        if (RGB) {   // six broadcast LDS.128 per item: keeps 24 registers free for the loads in flight
          const float4 r0 = wq[0], r1 = wq[1], q0 = wq[2], q1 = wq[3], b0 = wq[4], b1 = wq[5];
          w0[0] = r0.x; w0[1] = r0.y; w0[2] = r0.z; w0[3] = r0.w; w0[4] = r1.x; w0[5] = r1.y; w0[6] = r1.z; w0[7] = r1.w;
          w1[0] = q0.x; w1[1] = q0.y; w1[2] = q0.z; w1[3] = q0.w; w1[4] = q1.x; w1[5] = q1.y; w1[6] = q1.z; w1[7] = q1.w;
          w2[0] = b0.x; w2[1] = b0.y; w2[2] = b0.z; w2[3] = b0.w; w2[4] = b1.x; w2[5] = b1.y; w2[6] = b1.z; w2[7] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (RGB) {
            const float gt = w0[i] * ga[k] + w1[i] * gb[k] + w2[i] * gc[k];
            rrgb[i] = fmaf(ov[i], gt, rrgb[i]);
            rin[i] = fmaf(ov[i], gv[i], rin[i]);
            const float g = fmaf(sv[i], gt, gv[i] * si[i]);
            o[i] = g * kp[i] * (ov[i] > 0.f ? 1.f : 0.2f);
          } else {
            rin[i] = fmaf(ov[i], gv[i], rin[i]);
            o[i] = gv[i] * (ov[i] > 0.f ? si[i] * kp[i] : si[i] * kp[i] * 0.2f);
          }
          racc[i] = fmaf(o[i], nz[k] + bv[i], racc[i]);
        }
        stg8(a.gz + (img + p0 + p) * C + cv * 8, pack8(o));
      }
    }
    // gy*y = g*out - gy*(nz+b);  sum g*out = s_in*rin (+ s_rgb*rrgb);  racc was accumulated on gz = gy*d  (see act_bwd_kernel)
#pragma unroll
    for (int i = 0; i < 8; ++i) racc[i] = si[i] * rin[i] + (RGB ? sv[i] * rrgb[i] : 0.f) - racc[i] * SFK_SQRT2 / kp[i];
  }
  // ---- per-channel sums: block-level shared atomics, then one global atomic per channel
  auto flush = [&](const float* acc, float* gdst) {
    __syncthreads();
    for (int i = tid; i < C; i += kThreads) sacc[i] = 0.f;
    __syncthreads();
    if (tid < kConsumers) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sacc[cv * 8 + i], acc[i]);
    }
    __syncthreads();
    for (int i = tid; i < C; i += kThreads) atomicAdd(gdst + i, sacc[i]);
  };
  flush(racc, a.gdacc + static_cast<long>(n) * C);
  if (RGB) flush(rrgb, a.gs_rgb + static_cast<long>(n) * a.gs_stride);
  if (a.gs_in != nullptr) flush(rin, a.gs_in + static_cast<long>(n) * a.gin_stride);
}

// ---------------------------------------------------------------------------------------------
// ToRGB forward, streamed: rgb[n][c][p] = bias[c] + sum_i (wrgb[c][i] s[n][i]) x[n][p][i] + upsample2(skip)[n][c][p]
// for the two layers that carry 3/4 of its bytes (C = 32 at 1024^2, C = 64 at 512^2).
// One consumer thread owns one pixel of a 128-pixel chunk (the chunk is one piece of an image row), so there is no cross-lane
// reduction and a warp stores 128 contiguous bytes per colour plane.  The chunk sits in shared memory exactly as in HBM (rows of
// C*2 bytes), so lane l walks its pixel's 16-byte vectors in the rotated order v = (j + rot(l)) mod C/8, which puts the eight
// lanes of every LDS.128 phase on eight different bank groups; the matching weight vectors are [3][8] floats at a 112-byte pitch
// (28 banks: eight consecutive vectors never collide).  The 12 skip taps are fetched before the wait for the chunk.
constexpr int kTgPix = 128;                // pixels per chunk = consumer threads
constexpr int kTgThreads = kTgPix + 32;    // + producer warp
constexpr int kTgWPitch = 28;              // floats between the weight blocks of consecutive channel vectors
constexpr int kTgSkipCols = 72;            // skip columns staged per chunk: [w0/2 - 4, w0/2 + 68) covers every tap of 128 fine pixels
constexpr int kTgSkipBytes = 1792;         // 3 colours x 2 rows x 72 floats (1728 B), padded to a multiple of 128

struct TorgbStreamK {
  const __nv_bfloat16* x;
  const float* wrgb;    // [3][C]
  const float* s;       // + n * s_stride
  const float* bias;    // [3]
  const float* skip;    // [n][3][H/2][W/2] or null
  float* rgb;           // [n][3][H][W]
  int s_stride, H, W, C, chunks, chunk_bytes;
};

// upfirdn2d(skip, k*4, up=2, pad=(2,1)) at fine pixel (o,p): rows (i0, i0+1) x cols (j0, j0+1) with weights (a0, 1-a0) x (b0, 1-b0);
// even o -> (o/2-1: 1/4, o/2: 3/4); odd o -> ((o-1)/2: 3/4, (o+1)/2: 1/4).  The two skip rows of a chunk ride in the ring next to
// the activations (rows outside the image are not copied and are masked here).
__device__ __forceinline__ float skip_from_ring(uint32_t rows /* [2][kTgSkipCols] floats of one colour */, int jl /* j0 - first staged column */,
                                                bool vi0, bool vi1, bool vj0, bool vj1, float a0, float b0) {
  const float t00 = (vi0 && vj0) ? lds32(rows + jl * 4) : 0.f, t01 = (vi0 && vj1) ? lds32(rows + jl * 4 + 4) : 0.f;
  const float t10 = (vi1 && vj0) ? lds32(rows + (kTgSkipCols + jl) * 4) : 0.f, t11 = (vi1 && vj1) ? lds32(rows + (kTgSkipCols + jl) * 4 + 4) : 0.f;
  const float b1 = 1.f - b0;
  return a0 * (b0 * t00 + b1 * t01) + (1.f - a0) * (b0 * t10 + b1 * t11);
}

template <int VECS>   // C / 8
__global__ void __launch_bounds__(kTgThreads, 3) torgb_stream_kernel(const __grid_constant__ TorgbStreamK a) {
  extern __shared__ __align__(128) uint8_t sm_raw[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const uint32_t ring = (s_u32(sm_raw) + 127u) & ~127u;
  const int stage_bytes = a.chunk_bytes + kTgSkipBytes;
  float* const sw = reinterpret_cast<float*>(sm_raw + (ring - s_u32(sm_raw)) + kStages * stage_bytes);   // [VECS] blocks of [3][8], pitch 28 floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, C = VECS * 8;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mb_init(&full_bar[i], 1);
      mb_init(&empty_bar[i], kTgPix / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 3 * C; i += kTgThreads) {
    const int col = i / C, c = i % C;
    sw[(c >> 3) * kTgWPitch + col * 8 + (c & 7)] = a.wrgb[i] * a.s[static_cast<long>(n) * a.s_stride + c];
  }
  __syncthreads();
  const long HW = static_cast<long>(a.H) * a.W;
  if (warp == kTgPix / 32) {
    if (lane == 0) {
      int it = 0;
      for (int ch = blockIdx.x; ch < a.chunks; ch += gridDim.x, ++it) {
        const int st = it % kStages;
        mb_wait(&empty_bar[st], ((it / kStages) & 1) ^ 1);
        const uint32_t base = ring + st * stage_bytes;
        // skip rows i0, i0+1 of this image row, columns [w0/2 - 4, w0/2 + 68) clipped to the image (16-byte granules both ends)
        const int p0 = ch * kTgPix, h = p0 / a.W, w0 = p0 - h * a.W;
        const int hs = a.H / 2, ws = a.W / 2;
        const int i0 = (h & 1) ? (h - 1) / 2 : h / 2 - 1, jstart = w0 / 2 - 4;
        const int cs = jstart < 0 ? 0 : jstart, ce = jstart + kTgSkipCols > ws ? ws : jstart + kTgSkipCols;
        const int nrows = a.skip ? ((i0 >= 0 ? 1 : 0) + (i0 + 1 < hs ? 1 : 0)) : 0;
        mb_expect_tx(&full_bar[st], static_cast<uint32_t>(a.chunk_bytes + 3 * nrows * (ce - cs) * 4));
        bulk_g2s(base, a.x + (static_cast<long>(n) * HW + p0) * C, a.chunk_bytes, &full_bar[st]);
        if (a.skip) {
          for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 2; ++r) {
              const int i = i0 + r;
              if (i < 0 || i >= hs) continue;
              bulk_g2s(base + a.chunk_bytes + ((c * 2 + r) * kTgSkipCols + (cs - jstart)) * 4,
                       a.skip + ((static_cast<long>(n) * 3 + c) * hs + i) * ws + cs, (ce - cs) * 4, &full_bar[st]);
            }
        }
      }
    }
    return;
  }
  const int rot = VECS == 4 ? (lane >> 1) : lane;
  const uint32_t sw_u = s_u32(sw);
  const int hs = a.H / 2, ws = a.W / 2;
  const bool has_skip = a.skip != nullptr;
  float* const dst = a.rgb + static_cast<long>(n) * 3 * HW;
  const float bias0 = a.bias[0], bias1 = a.bias[1], bias2 = a.bias[2];
  int it = 0;
  for (int ch = blockIdx.x; ch < a.chunks; ch += gridDim.x, ++it) {
    const int st = it % kStages;
    const int p0 = ch * kTgPix;                 // a chunk lies inside one image row (W % 128 == 0)
    const int h = p0 / a.W, w0 = p0 - h * a.W, w = w0 + tid;
    const uint32_t px = ring + st * stage_bytes + tid * (VECS * 16);
    mb_wait(&full_bar[st], (it / kStages) & 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (has_skip) {      // the skip contribution first (its shared-memory reads precede the release of the slot)
      const int i0 = (h & 1) ? (h - 1) / 2 : h / 2 - 1, j0 = (w & 1) ? (w - 1) / 2 : w / 2 - 1;
      const bool vi0 = i0 >= 0, vi1 = i0 + 1 < hs, vj0 = j0 >= 0, vj1 = j0 + 1 < ws;
      const float fa = (h & 1) ? 0.75f : 0.25f, fb = (w & 1) ? 0.75f : 0.25f;
      const int jl = j0 - (w0 / 2 - 4);
      const uint32_t srow = ring + st * stage_bytes + a.chunk_bytes;
      a0 = skip_from_ring(srow, jl, vi0, vi1, vj0, vj1, fa, fb);
      a1 = skip_from_ring(srow + 2 * kTgSkipCols * 4, jl, vi0, vi1, vj0, vj1, fa, fb);
      a2 = skip_from_ring(srow + 4 * kTgSkipCols * 4, jl, vi0, vi1, vj0, vj1, fa, fb);
    }
#pragma unroll
    for (int j0 = 0; j0 < VECS; j0 += 4) {
      uint4 xr[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) xr[jj] = lds128(px + (((j0 + jj + rot) & (VECS - 1)) << 4));
      if (j0 + 4 >= VECS) {          // last read of the slot
        __syncwarp();
        if (lane == 0) mb_arrive(&empty_bar[st]);
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint32_t wp = sw_u + ((j0 + jj + rot) & (VECS - 1)) * (kTgWPitch * 4);
        const uint4 r0 = lds128(wp), r1 = lds128(wp + 16), q0 = lds128(wp + 32), q1 = lds128(wp + 48), b0 = lds128(wp + 64), b1 = lds128(wp + 80);
        float xv[8];
        unpack8(xr[jj], xv);
        const float wr[8] = {__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r0.w),
                             __uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z), __uint_as_float(r1.w)};
        const float wg[8] = {__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), __uint_as_float(q0.w),
                             __uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), __uint_as_float(q1.w)};
        const float wb[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                             __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a0 = fmaf(xv[i], wr[i], a0);
          a1 = fmaf(xv[i], wg[i], a1);
          a2 = fmaf(xv[i], wb[i], a2);
        }
      }
    }
    float r0 = a0 + bias0, r1 = a1 + bias1, r2 = a2 + bias2;
    float* o = dst + p0 + tid;
    o[0] = r0;
    o[HW] = r1;
    o[2 * HW] = r2;
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <bool RGB, bool GIN>
int launch(const ActStreamK& k, int n, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kStages) * kStageBytes + 128 + static_cast<size_t>(4 * k.C) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(act_stream_kernel<RGB, GIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) return static_cast<int>(e);
  int per_image = (2 * sfk_num_sms() + n - 1) / n;
  if (per_image > k.chunks) per_image = k.chunks;
  if (per_image < 1) per_image = 1;
  act_stream_kernel<RGB, GIN><<<dim3(static_cast<unsigned>(per_image), static_cast<unsigned>(n)), kThreads, smem, st>>>(k);
  return sfk_check_launch("act_stream");
}

}  // namespace

// Returns -1000 when the shape is not one the streaming form handles (the caller then runs the register kernel), otherwise the
// launch status.  wrgb == nullptr selects the plain activation backward (act_bwd), otherwise act_torgb_bwd.
int sfk_act_stream_launch(const void* out, const void* gin, void* gz, const float* d, const float* noise, float noise_w, const float* bias,
                          float* gdacc, const float* wrgb, const float* s_rgb, int s_stride, const float* grgb, float* gs_rgb, int gs_stride,
                          const float* s_in, float* gs_in, int in_stride, int gin_stride, int n, int hw, int c, cudaStream_t st) {
  static const int enabled = env_int("SFK_STREAM", 1);
  if (!enabled || sfk_act_f32()) return -1000;
  if (c < 32 || c > 512 || (c & (c - 1)) != 0) return -1000;
  const int P = kChunkBytes / (c * 2);
  if (P < 4 || hw % P != 0) return -1000;
  if (static_cast<long>(hw) * c * 2 < (1L << 20)) return -1000;   // small layers: the ring start-up is not amortised
  if (!sfk_aligned16(out) || !sfk_aligned16(gz) || (gin && !sfk_aligned16(gin)) || (grgb && !sfk_aligned16(grgb)) ||
      (noise && !sfk_aligned16(noise)) || (hw % 4) != 0)
    return -1000;
  ActStreamK k;
  k.out = static_cast<const __nv_bfloat16*>(out);
  k.gin = static_cast<const __nv_bfloat16*>(gin);
  k.gz = static_cast<__nv_bfloat16*>(gz);
  k.d = d; k.noise = noise; k.noise_w = noise_w; k.bias = bias; k.gdacc = gdacc;
  k.wrgb = wrgb; k.s_rgb = s_rgb; k.grgb = grgb; k.gs_rgb = gs_rgb; k.s_in = s_in; k.gs_in = gs_in;
  k.s_stride = s_stride; k.gs_stride = gs_stride; k.in_stride = in_stride; k.gin_stride = gin_stride;
  k.HW = hw; k.C = c; k.P = P; k.chunks = hw / P;
  for (k.vshift = 0; (8 << k.vshift) < c; ++k.vshift) {}
  if (wrgb != nullptr) return gin ? launch<true, true>(k, n, st) : launch<true, false>(k, n, st);
  if (gin == nullptr) return -1000;
  return launch<false, true>(k, n, st);
}

// ToRGB forward; same convention (-1000 = not handled)
int sfk_torgb_stream_launch(const void* x, const float* wrgb, const float* s, int s_stride, const float* bias, const float* skip, float* rgb,
                            int n, int h, int w, int c, cudaStream_t st) {
  static const int enabled = env_int("SFK_STREAM", 1);
  if (!enabled || sfk_act_f32()) return -1000;
  if (c != 32 && c != 64) return -1000;
  const long hw = static_cast<long>(h) * w;
  if (w % kTgPix != 0 || hw * c * 2 < (1L << 20) || !sfk_aligned16(x) || (h & 1)) return -1000;
  TorgbStreamK k;
  k.x = static_cast<const __nv_bfloat16*>(x);
  k.wrgb = wrgb; k.s = s; k.bias = bias; k.skip = skip; k.rgb = rgb;
  k.s_stride = s_stride; k.H = h; k.W = w; k.C = c; k.chunks = static_cast<int>(hw / kTgPix); k.chunk_bytes = kTgPix * c * 2;
  const size_t smem = static_cast<size_t>(kStages) * (k.chunk_bytes + kTgSkipBytes) + 128 + static_cast<size_t>(c / 8) * kTgWPitch * sizeof(float);
  const void* fn = c == 32 ? reinterpret_cast<const void*>(&torgb_stream_kernel<4>) : reinterpret_cast<const void*>(&torgb_stream_kernel<8>);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) return static_cast<int>(e);
  int per_image = (3 * sfk_num_sms() + n - 1) / n;
  if (per_image > k.chunks) per_image = k.chunks;
  void* args[1] = {&k};
  e = cudaLaunchKernel(fn, dim3(static_cast<unsigned>(per_image), static_cast<unsigned>(n)), dim3(kTgThreads), args, smem, st);
  if (e != cudaSuccess) {
    sfk_set_error(cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return sfk_check_launch("torgb_stream");
}
