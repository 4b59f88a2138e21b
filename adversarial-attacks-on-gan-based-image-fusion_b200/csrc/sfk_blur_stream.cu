// Streaming form of the blur of the unfused up-layers (upfirdn2d [1,3,3,1], pad (1,1), over the phase-planar transposed-conv
// output, fused with demodulation / noise / bias / leaky-ReLU) and of its transpose.
//
// The tiled kernel (sfk_elementwise.cu: blur_tile_kernel) alternates a load phase and a compute phase behind __syncthreads and
// spends 14 FMAs + 4.4 bf16 unpacks per output element on a 2x4 register block: measured 1.2 TB/s at 256^2 x 128 channels.
// Here a thread owns ONE coarse column (two fine columns) x 4 channels and marches down the rows of its strip
// (4 channels, not 8: the four pending rows then cost 32 registers and two CTAs of 8+1 warps stay resident per SM):
//   * a producer warp streams one fine input row per ring stage into shared memory with cp.async.bulk (both column phases of the
//     row, plus the noise row the stage's output needs), the consumer warps read it with conflict-free LDS.64;
//   * the FIR is separable: the 4-tap horizontal pass runs once per input row, the vertical pass is a scatter of that row into the
//     four pending output rows it touches (weights 1/4, 3/4, 3/4, 1/4); the row that became complete is finished and stored.  The
//     four pending rows live in registers and rotate by renaming (loop unrolled by four);
//   * all arithmetic is packed fp32 (FFMA2: two channels per instruction): 4 FMA issues + 2.5 unpack ops per output element.
// Out-of-range columns/rows are never copied: their ring slots are zeroed once, so the hot loop carries no bounds checks.
//
// Backward additionally finishes the data gradient that arrives from the next conv's plain dgrad launch (s_in, gs_in: multiply by
// the style, accumulate sum x*gx~ -- the same consumer-side finishing sfk_act_bwd does).
#include <stdlib.h>

#include <type_traits>

#include "sfk_common.cuh"

namespace {

constexpr int kMaxStages = 12;   // ring depth is chosen per launch: as many stages as the shared-memory budget holds
constexpr int kMaxConsumers = 512;   // 256 consumer threads (two CTAs per SM) for C <= 128, 512 (one CTA) above

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
// the barrier helpers take shared-window addresses formed ONCE per thread (forming one costs an S2R in the hot loop otherwise)
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mb_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a time slice): issued one step ahead, so that the latency of the barrier
// read overlaps the arithmetic of the current row instead of heading every step (22 % of the stall samples of the first version sat
// on the branch behind try_wait, profiles/blur_stream_r2.md)
__device__ __forceinline__ bool mb_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mb_wait(uint32_t bar, uint32_t parity) {   // bounded: a stuck ring traps instead of hanging the GPU
#pragma unroll 1
  for (long i = 0; i < (1L << 28); ++i)
    if (mb_try(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// 4 bf16 -> two channel pairs in fp32
struct F4 {
  float2 p[2];
};
__device__ __forceinline__ float2 up2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ F4 unpack(const uint2& v) {
  F4 r;
  r.p[0] = up2(v.x); r.p[1] = up2(v.y);
  return r;
}
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

struct BlurStreamK {
  const __nv_bfloat16* src0;   // fwd: T [n][4][H+1][W+1][C]          bwd: out [n][2H][2W][C]
  const __nv_bfloat16* src1;   //                                      bwd: gout
  __nv_bfloat16* dst;          // fwd: out                             bwd: gT (every slot of every plane is written)
  const float* d;              // [n][C]
  const float* noise;          // [2H][2W] or null
  float noise_w;
  const float* bias;           // [C]
  float* gdacc;                // bwd: [n][C] += sum gy*y
  const float* s_in;           // bwd: + n * in_stride, or null (gradient arrives finished)
  float* gs_in;                // bwd: + n * in_stride, or null
  int in_stride;
  int H, W, C, TJ, TI, strips, rblocks, vshift;
  int chunk_bytes, stage_bytes, stages;
};

// One FIR step: the horizontally filtered row h[2] (column phase b = 0, 1) is scattered into the pending rows
//   A (+1/4, complete afterwards), B (+3/4), Cc (+3/4), D (= 1/4, new)
__device__ __forceinline__ void scatter(const F4 (&h)[2], F4 (&A)[2], F4 (&B)[2], F4 (&Cc)[2], F4 (&D)[2]) {
  const float2 q = bc2(0.25f), t = bc2(0.75f);
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      A[b].p[i] = __ffma2_rn(q, h[b].p[i], A[b].p[i]);
      B[b].p[i] = __ffma2_rn(t, h[b].p[i], B[b].p[i]);
      Cc[b].p[i] = __ffma2_rn(t, h[b].p[i], Cc[b].p[i]);
      D[b].p[i] = __fmul2_rn(q, h[b].p[i]);
    }
}
__device__ __forceinline__ void clear(F4 (&D)[2]) {   // an input row outside the tensor contributes nothing
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int i = 0; i < 2; ++i) D[b].p[i] = make_float2(0.f, 0.f);
}

// MINB: CTAs per SM the register budget is planned for (3: forward with 256 consumers, 2: backward with 256, 1: 512 consumers)
template <bool BWD, int MINB>
__global__ void __launch_bounds__(MINB == 1 ? kMaxConsumers + 32 : 288, MINB) blur_stream_kernel(const __grid_constant__ BlurStreamK a) {
  extern __shared__ __align__(128) uint8_t sm_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  const int kStages = a.stages;
  const uint32_t full0 = s_u32(&full_bar[0]), empty0 = s_u32(&empty_bar[0]);
  const uint32_t ring = (s_u32(sm_raw) + 127u) & ~127u;
  uint8_t* const ring_p = sm_raw + (ring - s_u32(sm_raw));
  float* const sacc = reinterpret_cast<float*>(ring_p + kStages * a.stage_bytes);   // [C] (backward reductions)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = blockDim.x, consumers = nthreads - 32;
  const int H = a.H, W = a.W, C = a.C, TJ = a.TJ;
  int item = blockIdx.x;
  const int rb = item % a.rblocks;
  item /= a.rblocks;
  const int strip = item % a.strips, n = item / a.strips;
  const int j0 = strip * TJ, i0 = rb * a.TI;
  // input rows streamed by this CTA and the range of output rows it owns
  //   fwd: T rows q in [2 i0 - 1, 2 (i0+TI) + 1], output rows [2 i0, 2 (i0+TI)) (< 2H); row q completes output row q - 2
  //   bwd: gradient rows o in [2 i0 - 2, 2 (i0+TI)], T rows [2 i0, 2 (i0+TI)) (<= 2H+1); row o completes T row o - 1
  const int r_lo = BWD ? 2 * i0 - 2 : 2 * i0 - 1, r_hi = BWD ? 2 * (i0 + a.TI) : 2 * (i0 + a.TI) + 1;
  const int e_off = BWD ? 1 : 2;
  const int e_lo = 2 * i0, e_hi = min(2 * (i0 + a.TI), BWD ? 2 * H + 2 : 2 * H);
  const int in_rows = BWD ? 2 * H : 2 * H + 1;   // valid input rows [0, in_rows)
  // columns of the ring chunk: slot s holds  fwd: plane column j0 - 1 + s (TJ + 2 slots)   bwd: pixel 2 j0 - 2 + s (2 TJ + 3 slots)
  const int nslots = BWD ? 2 * TJ + 3 : TJ + 2;
  const int c_first = BWD ? 2 * j0 - 2 : j0 - 1;
  const int c_lo = max(c_first, 0);
  const int c_hi0 = min(c_first + nslots, BWD ? 2 * W : W + 1);   // chunk 0 (fwd: even columns, W+1 of them)
  const int c_hi1 = min(c_first + nslots, BWD ? 2 * W : W);       // chunk 1 (fwd: odd columns, W valid)
  const int px_bytes = C * 2;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mb_init(&full_bar[i], 1);
      mb_init(&empty_bar[i], consumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the slots no copy ever writes (columns outside the tensor)
  {
    const int vec_per_px = px_bytes / 16;
    for (int e = tid; e < kStages * 2 * nslots * vec_per_px; e += nthreads) {
      const int v = e % vec_per_px, s = (e / vec_per_px) % nslots, ch = (e / (vec_per_px * nslots)) % 2, st = e / (vec_per_px * nslots * 2);
      const int col = c_first + s;
      if (col < c_lo || col >= (ch ? c_hi1 : c_hi0))
        *reinterpret_cast<uint4*>(ring_p + st * a.stage_bytes + ch * a.chunk_bytes + s * px_bytes + v * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();

  const int vecs = C >> 2;                       // 4-channel vectors per pixel
  const int cv = tid & (vecs - 1), col = tid >> a.vshift;
  const int j = j0 + col;
  const bool active = tid < consumers && col < TJ;
  // backward reductions over this thread's own pixels: rin = sum out*gout, rg = sum g', rgn = sum g' * noise  (g' = gy * d)
  float rin[4], rg[4], rgn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) rin[i] = rg[i] = rgn[i] = 0.f;
  float2 rin2[2], rg2[2], rgn2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) rin2[i] = rg2[i] = rgn2[i] = make_float2(0.f, 0.f);

  if (warp == consumers / 32) {
    // ===================== producer =====================
    if (lane == 0) {
      const int Hp = H + 1, Wp = W + 1;
      // noise columns [2 j0, 2 j0 + 2 TJ) of the row this stage's consumer needs (fwd: the completed output row; bwd: the input row)
      const int nz_c0 = 2 * j0, nz_c1 = min(2 * j0 + 2 * TJ, 2 * W);
      int st = 0;
      uint32_t ph = 0;
      for (int r = r_lo; r <= r_hi; ++r, ++st) {
        if (st == kStages) {
          st = 0;
          ph ^= 1u;
        }
        if (!mb_try(empty0 + st * 8, ph ^ 1u)) mb_wait(empty0 + st * 8, ph ^ 1u);
        const uint32_t fb = full0 + st * 8;
        const uint32_t base = ring + st * a.stage_bytes;
        const bool row_ok = r >= 0 && r < in_rows;
        const int nr = BWD ? r : r - e_off;    // noise row
        const bool nz_ok = a.noise != nullptr && nz_c1 > nz_c0 && nr >= 0 && nr < 2 * H && (BWD || (nr >= e_lo && nr < e_hi));
        uint32_t tx = 0;
        if (row_ok) tx += static_cast<uint32_t>((c_hi0 - c_lo) + (c_hi1 - c_lo)) * px_bytes;
        if (nz_ok) tx += static_cast<uint32_t>(nz_c1 - nz_c0) * 4;
        mb_expect_tx(fb, tx);
        if (row_ok) {
          const uint32_t doff = static_cast<uint32_t>(c_lo - c_first) * px_bytes;
          if (BWD) {
            const long off = ((static_cast<long>(n) * 2 * H + r) * 2 * W + c_lo) * C;
            bulk_g2s(base + doff, a.src0 + off, static_cast<uint32_t>(c_hi0 - c_lo) * px_bytes, fb);
            bulk_g2s(base + a.chunk_bytes + doff, a.src1 + off, static_cast<uint32_t>(c_hi1 - c_lo) * px_bytes, fb);
          } else {
            const long p0 = ((static_cast<long>(n) * 4 + (r & 1) * 2) * Hp + (r >> 1)) * Wp;      // even-column plane of this row
            const long p1 = p0 + static_cast<long>(Hp) * Wp;                                       // odd-column plane
            bulk_g2s(base + doff, a.src0 + (p0 + c_lo) * C, static_cast<uint32_t>(c_hi0 - c_lo) * px_bytes, fb);
            bulk_g2s(base + a.chunk_bytes + doff, a.src0 + (p1 + c_lo) * C, static_cast<uint32_t>(c_hi1 - c_lo) * px_bytes, fb);
          }
        }
        if (nz_ok) bulk_g2s(base + 2 * a.chunk_bytes, a.noise + static_cast<long>(nr) * 2 * W + nz_c0, static_cast<uint32_t>(nz_c1 - nz_c0) * 4, fb);
      }
    }
  } else {
    // ===================== consumers =====================
    float2 dv[2], bv[2], kp[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) dv[i] = bv[i] = kp[i] = make_float2(0.f, 0.f);
    if (active) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = cv * 4 + 2 * i;
        const float d0 = a.d[static_cast<long>(n) * C + c], d1 = a.d[static_cast<long>(n) * C + c + 1];
        if (BWD) {   // g' = gout * s_in * act'(out) * d  (d rides on the gradient: the FIR is linear)
          const float s0 = a.s_in ? a.s_in[static_cast<long>(n) * a.in_stride + c] : 1.f, s1 = a.s_in ? a.s_in[static_cast<long>(n) * a.in_stride + c + 1] : 1.f;
          kp[i] = make_float2(s0 * d0 * SFK_SQRT2, s1 * d1 * SFK_SQRT2);
        } else {     // out = max(u, 0.2 u),  u = sqrt2 * (acc * d + noise + bias)
          dv[i] = make_float2(d0 * SFK_SQRT2, d1 * SFK_SQRT2);
          bv[i] = make_float2(a.bias[c] * SFK_SQRT2, a.bias[c + 1] * SFK_SQRT2);
        }
      }
    }
    const float nw = BWD ? a.noise_w : a.noise_w * SFK_SQRT2;
    const bool has_noise = a.noise != nullptr;
    const bool own = j < W;                     // bwd: this thread's two fine pixels exist (the last strip holds only column W)
    const bool writer = active && (!BWD || j <= W);
    const bool edge = BWD && j == W;            // bwd: T column 2W+1 does not exist (its slot is written as zero)
    const int Hp = H + 1, Wp = W + 1;
    const uint32_t sb = static_cast<uint32_t>(a.stage_bytes), cb = static_cast<uint32_t>(a.chunk_bytes), pb = static_cast<uint32_t>(px_bytes);
    const uint32_t my = static_cast<uint32_t>((BWD ? 2 * col : col) * px_bytes + cv * 8);
    const uint32_t nz_off = 2 * cb + static_cast<uint32_t>(col * 8);
    // element strides of the destination rows
    const long fwd_row = static_cast<long>(2 * W) * C;
    __nv_bfloat16* const fwd_base = a.dst + (static_cast<long>(n) * 2 * H * 2 * W + 2 * j) * C + cv * 4;                 // + er * fwd_row
    __nv_bfloat16* const bwd_base = a.dst + ((static_cast<long>(n) * 4 * Hp) * Wp + j) * C + cv * 4;                     // + ((plane * Hp + m) * Wp) * C
    F4 S0[2], S1[2], S2[2], S3[2];
    clear(S0); clear(S1); clear(S2); clear(S3);
    int st = 0;
    uint32_t ph = 0, base = ring;
    bool pre_ok = false;
    // FULL: the row is inside the tensor, the completed row is one of this CTA's, reductions are this CTA's (steady state of the march)
    auto step = [&](int r, auto full_tag, F4 (&A)[2], F4 (&B)[2], F4 (&Cc)[2], F4 (&D)[2]) {
      constexpr bool FULL = decltype(full_tag)::value;
      if (!pre_ok && !mb_try(full0 + st * 8, ph)) mb_wait(full0 + st * 8, ph);
      const uint32_t sbase = base;
      const int er = r - e_off;                  // row completed by this step
      const bool row_ok = FULL || (r >= 0 && r < in_rows);
      const bool emit = FULL || (er >= e_lo && er < e_hi);
      const bool red = own && (FULL || (r >= e_lo && r < 2 * (i0 + a.TI)));   // halo rows belong to the neighbouring row block's reductions
      float2 nz = make_float2(0.f, 0.f);
      if (row_ok && active) {
        F4 h[2];
        const float2 q = bc2(0.25f), t = bc2(0.75f);
        if (BWD) {
          // five fine pixels 2j-2 .. 2j+2 of the gradient row: g' = gout * (out > 0 ? kp : 0.2 kp); filtered on the fly:
          //   T column 2j   = 1/4 g(2j-2) + 3/4 g(2j-1) + 3/4 g(2j) + 1/4 g(2j+1);   2j+1: the same one pixel to the right
          if (has_noise) nz = lds64f(sbase + nz_off);
          uint2 ov[5], gv[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            ov[k] = lds64(sbase + my + k * pb);
            gv[k] = lds64(sbase + cb + my + k * pb);
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const F4 o = unpack(ov[k]), gg = unpack(gv[k]);
            F4 g;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float2 kn = __fmul2_rn(kp[i], bc2(0.2f));
              const float2 m = make_float2(o.p[i].x > 0.f ? kp[i].x : kn.x, o.p[i].y > 0.f ? kp[i].y : kn.y);
              g.p[i] = __fmul2_rn(gg.p[i], m);
              if (k == 0) h[0].p[i] = __fmul2_rn(q, g.p[i]);
              if (k == 1) { h[0].p[i] = __ffma2_rn(t, g.p[i], h[0].p[i]); h[1].p[i] = __fmul2_rn(q, g.p[i]); }
              if (k == 2) { h[0].p[i] = __ffma2_rn(t, g.p[i], h[0].p[i]); h[1].p[i] = __ffma2_rn(t, g.p[i], h[1].p[i]); }
              if (k == 3) { h[0].p[i] = __ffma2_rn(q, g.p[i], h[0].p[i]); h[1].p[i] = __ffma2_rn(t, g.p[i], h[1].p[i]); }
              if (k == 4) h[1].p[i] = __ffma2_rn(q, g.p[i], h[1].p[i]);
            }
            if ((k == 2 || k == 3) && red) {     // own pixels 2j, 2j+1: style gradient sum x*gx~ and the demodulation sums
              const float2 nzk = bc2(k == 2 ? nz.x : nz.y);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                rin2[i] = __ffma2_rn(o.p[i], gg.p[i], rin2[i]);
                rg2[i] = __fadd2_rn(rg2[i], g.p[i]);
                rgn2[i] = __ffma2_rn(g.p[i], nzk, rgn2[i]);
              }
            }
          }
        } else {
          // even-column plane at j, j+1 (slots col+1, col+2), odd-column plane at j-1, j, j+1 (slots col, col+1, col+2)
          const uint2 re0 = lds64(sbase + my + pb), re1 = lds64(sbase + my + 2 * pb);
          const uint2 rom = lds64(sbase + cb + my), ro0 = lds64(sbase + cb + my + pb), ro1 = lds64(sbase + cb + my + 2 * pb);
          if (emit && has_noise) nz = lds64f(sbase + nz_off);
          const F4 e0 = unpack(re0), e1 = unpack(re1), om = unpack(rom), o0 = unpack(ro0), o1 = unpack(ro1);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            // fine column 2j:   1/4 T(2j-1) + 3/4 T(2j) + 3/4 T(2j+1) + 1/4 T(2j+2);   2j+1: shifted by one
            h[0].p[i] = __ffma2_rn(q, e1.p[i], __ffma2_rn(t, o0.p[i], __ffma2_rn(t, e0.p[i], __fmul2_rn(q, om.p[i]))));
            h[1].p[i] = __ffma2_rn(q, o1.p[i], __ffma2_rn(t, e1.p[i], __ffma2_rn(t, o0.p[i], __fmul2_rn(q, e0.p[i]))));
          }
        }
        scatter(h, A, B, Cc, D);
      } else {
        clear(D);
        if (!BWD && emit && has_noise && active) nz = lds64f(sbase + nz_off);
      }
      __syncwarp();
      if (lane == 0) mb_arrive(empty0 + st * 8);   // every shared-memory read of the stage is done
      base += sb;
      if (++st == kStages) {
        st = 0;
        ph ^= 1u;
        base = ring;
      }
      pre_ok = mb_test(full0 + st * 8, ph);      // the next row's barrier, read while this row is finished and stored
      if (emit && writer) {
        if (BWD) {
          const bool zrow = !FULL && er > 2 * H;   // row 2H+1 does not exist either
          __nv_bfloat16* const o0 = bwd_base + (static_cast<long>((er & 1) * 2 * Hp + (er >> 1)) * Wp) * C;
          const long plane = static_cast<long>(Hp) * Wp * C;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const bool zero = zrow || (b == 1 && edge);
            uint2 o;
            o.x = zero ? 0u : pack2(A[b].p[0].x, A[b].p[0].y);
            o.y = zero ? 0u : pack2(A[b].p[1].x, A[b].p[1].y);
            *reinterpret_cast<uint2*>(o0 + b * plane) = o;
          }
        } else {
          // the two fine pixels of this thread are neighbours in memory: out[er][2j + b][cv*4 ..]
          __nv_bfloat16* const o0 = fwd_base + er * fwd_row;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const float2 nb = bc2(nw * (b == 0 ? nz.x : nz.y));
            const float2 u0 = __ffma2_rn(A[b].p[0], dv[0], __fadd2_rn(nb, bv[0])), u1 = __ffma2_rn(A[b].p[1], dv[1], __fadd2_rn(nb, bv[1]));
            const float2 l0 = __fmul2_rn(u0, bc2(0.2f)), l1 = __fmul2_rn(u1, bc2(0.2f));
            uint2 w;
            w.x = pack2(fmaxf(u0.x, l0.x), fmaxf(u0.y, l0.y));
            w.y = pack2(fmaxf(u1.x, l1.x), fmaxf(u1.y, l1.y));
            *reinterpret_cast<uint2*>(o0 + b * C) = w;
          }
        }
      }
    };
    auto step_role = [&](int r, int role, auto tag) {
      switch (role) {
        case 0: step(r, tag, S0, S1, S2, S3); break;
        case 1: step(r, tag, S1, S2, S3, S0); break;
        case 2: step(r, tag, S2, S3, S0, S1); break;
        default: step(r, tag, S3, S0, S1, S2); break;
      }
    };
    // steady-state rows [rs, re): inside the tensor, completing one of this CTA's rows, reductions owned by this CTA
    int rs = max(max(0, e_lo + e_off), r_lo), re = min(min(in_rows, e_hi + e_off), r_hi + 1);
    if (BWD) re = min(re, 2 * (i0 + a.TI));
    rs += (4 - ((rs - r_lo) & 3)) & 3;           // role 0 at the start of the unrolled loop
    int r = r_lo;
    for (; r <= r_hi && r < rs; ++r) step_role(r, (r - r_lo) & 3, std::false_type{});
    for (; r + 3 < re; r += 4) {
      step(r, std::true_type{}, S0, S1, S2, S3);
      step(r + 1, std::true_type{}, S1, S2, S3, S0);
      step(r + 2, std::true_type{}, S2, S3, S0, S1);
      step(r + 3, std::true_type{}, S3, S0, S1, S2);
    }
    for (; r <= r_hi; ++r) step_role(r, (r - r_lo) & 3, std::false_type{});
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      rin[2 * i] = rin2[i].x; rin[2 * i + 1] = rin2[i].y;
      rg[2 * i] = rg2[i].x; rg[2 * i + 1] = rg2[i].y;
      rgn[2 * i] = rgn2[i].x; rgn[2 * i + 1] = rgn2[i].y;
    }
  }
  if (BWD) {
    // sum gy*y = sum g*out - sum gy*(noise + bias) = s_in * rin - (noise_w * rgn + bias * rg) / d     (rg, rgn were taken on g' = gy*d)
    float racc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) racc[i] = 0.f;
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = cv * 4 + i;
        const float si = a.s_in ? a.s_in[static_cast<long>(n) * a.in_stride + c] : 1.f;
        racc[i] = si * rin[i] - (a.noise_w * rgn[i] + a.bias[c] * rg[i]) / a.d[static_cast<long>(n) * C + c];
      }
    }
    auto flush = [&](const float* acc, float* gdst) {
      __syncthreads();
      for (int i = tid; i < C; i += nthreads) sacc[i] = 0.f;
      __syncthreads();
      if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(&sacc[cv * 4 + i], acc[i]);
      }
      __syncthreads();
      for (int i = tid; i < C; i += nthreads) atomicAdd(gdst + i, sacc[i]);
    };
    flush(racc, a.gdacc + static_cast<long>(n) * C);
    if (a.gs_in != nullptr) flush(rin, a.gs_in + static_cast<long>(n) * a.in_stride);
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// Returns -1000 when the shape is outside the streaming kernel's domain (the caller then runs the tiled kernel).
int sfk_blur_stream_launch(bool bwd, const void* a0, const void* a1, void* dst, const float* d, const float* noise, float noise_w, const float* bias,
                           float* gdacc, const float* s_in, float* gs_in, int in_stride, int n, int h, int w, int c, cudaStream_t st) {
  static const int enabled = env_int("SFK_BLUR_STREAM", 1);
  if (!enabled || sfk_act_f32()) return -1000;
  if (c < 32 || c > 512 || (c & (c - 1)) != 0 || h < 2 || w < 2) return -1000;
  if (!sfk_aligned16(a0) || (a1 && !sfk_aligned16(a1)) || !sfk_aligned16(dst) || (noise && !sfk_aligned16(noise))) return -1000;
  const int vecs = c / 4;
  const int consumers = c <= 128 ? 256 : kMaxConsumers;
  int TJ = consumers / vecs;
  if (TJ > w) TJ = w;
  if (TJ < 2 || w % TJ != 0) return -1000;
  BlurStreamK k;
  k.src0 = static_cast<const __nv_bfloat16*>(a0);
  k.src1 = static_cast<const __nv_bfloat16*>(a1);
  k.dst = static_cast<__nv_bfloat16*>(dst);
  k.d = d; k.noise = noise; k.noise_w = noise_w; k.bias = bias; k.gdacc = gdacc; k.s_in = s_in; k.gs_in = gs_in; k.in_stride = in_stride;
  k.H = h; k.W = w; k.C = c; k.TJ = TJ;
  k.strips = bwd ? w / TJ + 1 : w / TJ;
  for (k.vshift = 0; (4 << k.vshift) < c; ++k.vshift) {}
  const int nslots = bwd ? 2 * TJ + 3 : TJ + 2;
  k.chunk_bytes = nslots * c * 2;
  k.stage_bytes = ((2 * k.chunk_bytes + 2 * TJ * 4 + 127) / 128) * 128;
  // row blocks: one wave of CTAs where the layer is large enough, at least 8 coarse rows per block
  const int rows = bwd ? h + 1 : h;
  const int per_sm = consumers == 256 ? (bwd ? 2 : 3) : 1;
  const int slots = per_sm * sfk_num_sms();
  int rblocks = slots / (n * k.strips);
  if (rblocks < 1) rblocks = 1;
  if (rblocks > (rows + 7) / 8) rblocks = (rows + 7) / 8;
  k.TI = (rows + rblocks - 1) / rblocks;
  k.rblocks = (rows + k.TI - 1) / k.TI;
  const int budget = (per_sm == 3 ? 68 : per_sm == 2 ? 100 : 200) * 1024 - 128 - c * static_cast<int>(sizeof(float));
  k.stages = budget / k.stage_bytes;
  if (k.stages > kMaxStages) k.stages = kMaxStages;
  if (k.stages < 3) return -1000;
  const size_t smem = static_cast<size_t>(k.stages) * k.stage_bytes + 128 + static_cast<size_t>(c) * sizeof(float);
  const void* fn = per_sm == 1 ? (bwd ? reinterpret_cast<const void*>(&blur_stream_kernel<true, 1>) : reinterpret_cast<const void*>(&blur_stream_kernel<false, 1>))
                               : (bwd ? reinterpret_cast<const void*>(&blur_stream_kernel<true, 2>) : reinterpret_cast<const void*>(&blur_stream_kernel<false, 3>));
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return static_cast<int>(e);
  void* args[1] = {&k};
  e = cudaLaunchKernel(fn, dim3(static_cast<unsigned>(n * k.strips * k.rblocks)), dim3(consumers + 32), args, smem, st);
  if (e != cudaSuccess) {
    sfk_set_error(cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return sfk_check_launch(bwd ? "blur_stream_bwd" : "blur_stream_fwd");
}
