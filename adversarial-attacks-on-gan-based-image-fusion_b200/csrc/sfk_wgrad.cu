// Weight-gradient GEMM of the 3x3 convolutions (the second backward GEMM of conv3x3+bias+ReLU and of ModulatedConv2d):
//     dw[s][tap][co][ci] += sum_{n in s} sum_{h,w} gz[n][h][w][co] * x[n][h+dy][w+dx][ci]          (tap = (dy+1)*3 + dx+1)
// on the tensor cores.  The reference never freezes parameters (attack_main2.py:301-304), so its autograd runs this GEMM for every
// conv on every iteration; the attack itself does not need it (the engines keep the weights frozen), which is why it is a
// stand-alone operator here and not a stage of the engines' backward.
//
// GEMM view: M = 128 output channels (rows of dw), N = 64/128 input channels, K = pixels.  Both operands are the NHWC activation
// tensors as they lie in HBM: a TMA box {64 channels, TW, TH} lands in shared memory as one 128-byte row per pixel, i.e. as an
// MN-major operand (channels contiguous) whose K index walks the pixel rows: UMMA canonical layout
//     Swizzle<3,4,3> o ((8,m),(8,k)) : ((1,LBO),(8,SBO))   in 16-byte units
// with SBO = 1024 B (the next 8 pixels) and LBO = the distance between the boxes of two 64-channel groups.  No transposition, no
// im2col: the filter tap (dy,dx) is a coordinate offset of the x box (dx) and a row offset of the B operand inside it (dy), the
// conv padding is TMA's zero fill.  The 9 taps need 9 accumulators of N columns; TMEM holds 512, so a CTA owns the three taps of one
// dx (3 x 128 columns) and the pixel tiles of one split; splits are reduced with fp32 atomics (red.global.add).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2-5 = epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "sfk_common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kTilePx = 64;   // pixels (K) per pipeline stage

struct __align__(64) WgradArgs {
  CUtensorMap mapA;   // gz: {cout, W, H, N}, box {64, TW, TH, 1}
  CUtensorMap mapB;   // x:  {cin,  W, H, N}, box {64, TW, TH + 2, 1}
  float* dw;          // [S][9][cout][cin]
  int n_img, h, w, cout, cin, per_sample;
  int TH, TW, tiles_h, tiles_w;
  int m_boxes, n_boxes;            // 64-channel boxes per operand (M = 64 * m_boxes is always issued as 128: rows beyond cout are zero)
  int ci_blocks, splits, stages;
  int a_box_bytes, b_box_bytes, stage_bytes;
  int* err;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded: a pipeline that never advances raises *err and lets the kernel drain instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  for (long i = 0; i < (1L << 26); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((i & 1023) == 1023 && err && *reinterpret_cast<volatile int*>(err) != 0) break;
  }
  if (err) atomicExch(err, 1);
  return false;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// MN-major SW128 operand descriptor: start address, LBO (between 64-element MN groups), SBO (between 8-row K groups)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1u) << 46;   // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2u) << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // work item: (output-channel block, input-channel block, dx) x split
  int item = blockIdx.y;
  const int dxi = item % 3;
  item /= 3;
  const int cib = item % a.ci_blocks, cob = item / a.ci_blocks;
  const int co0 = cob * 128, ci0 = cib * 64 * a.n_boxes;
  const int Nb = 64 * a.n_boxes;
  const int split = blockIdx.x;
  const int tiles_img = a.tiles_h * a.tiles_w;
  // shared weights: the tiles of all images, strided over the splits; per-sample: `splits` = images x splits-per-image
  int t_first, t_step, t_end, img_fixed;
  if (a.per_sample) {
    const int spi = a.splits / a.n_img;
    img_fixed = split / spi;
    t_first = split % spi; t_step = spi; t_end = tiles_img;
  } else {
    img_fixed = -1;
    t_first = split; t_step = a.splits; t_end = tiles_img * a.n_img;
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapB)) : "memory");
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(a.m_boxes * a.a_box_bytes + a.n_boxes * a.b_box_bytes);
      int it = 0;
      for (int t = t_first; t < t_end; t += t_step, ++it) {
        const int st = it % a.stages;
        if (!mbar_wait(&empty_bar[st], ((it / a.stages) & 1) ^ 1, a.err)) break;
        const int img = a.per_sample ? img_fixed : t / tiles_img, tt = a.per_sample ? t : t % tiles_img;
        const int h0 = (tt / a.tiles_w) * a.TH, w0 = (tt % a.tiles_w) * a.TW;
        const uint32_t sa = smem_base + st * a.stage_bytes, sb = sa + a.m_boxes * a.a_box_bytes;
        mbar_expect_tx(&full_bar[st], tx);
        for (int mb = 0; mb < a.m_boxes; ++mb) tma_load_4d(sa + mb * a.a_box_bytes, &a.mapA, &full_bar[st], co0 + 64 * mb, w0, h0, img);
        for (int nb = 0; nb < a.n_boxes; ++nb)
          tma_load_4d(sb + nb * a.b_box_bytes, &a.mapB, &full_bar[st], ci0 + 64 * nb, w0 + dxi - 1, h0 - 1, img);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(Nb >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);
      // M = 128 is always issued; with one 64-channel box the second MN group re-reads the first (LBO = 0): its rows are channels
      // >= cout of dw and are never stored
      const uint32_t lbo_a = a.m_boxes == 2 ? a.a_box_bytes : 0u, lbo_b = a.n_boxes == 2 ? a.b_box_bytes : 0u;
      int it = 0;
      bool ok = true;
      for (int t = t_first; t < t_end && ok; t += t_step, ++it) {
        const int st = it % a.stages;
        ok = mbar_wait(&full_bar[st], (it / a.stages) & 1, a.err);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + st * a.stage_bytes, sb = sa + a.m_boxes * a.a_box_bytes;
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll
          for (int ks = 0; ks < kTilePx / 16; ++ks) {
            const uint64_t ad = make_desc_mn(sa + ks * 2048, lbo_a, 1024);
            const uint64_t bd = make_desc_mn(sb + (dyi * a.TW + ks * 16) * 128, lbo_b, 1024);
            umma_bf16(tmem_base + static_cast<uint32_t>(dyi * Nb), ad, bd, idesc, (it == 0 && ks == 0) ? 0u : 1u);
          }
        }
        umma_commit(&empty_bar[st]);
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // ===================== epilogue: TMEM -> red.global.add =====================
    const bool any = t_first < t_end;
    const bool ok = mbar_wait(&tmem_full_bar, 0, a.err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                    // TMEM lane quarter this warp may read (warps 2..5 -> 2,3,0,1)
    const int row = q * 32 + lane, co = co0 + row;
    const int s = a.per_sample ? img_fixed : 0;
    if (ok && any) {
      for (int dyi = 0; dyi < 3; ++dyi) {
        const int tap = dyi * 3 + dxi;
        float* const drow = a.dw + ((static_cast<long>(s) * 9 + tap) * a.cout + co) * a.cin + ci0;
        for (int c0 = 0; c0 < Nb; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(dyi * Nb + c0), v);
          if (co < a.cout) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ci0 + c0 + i < a.cin) atomicAdd(drow + c0 + i, v[i]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// CUDA-core form with the identical contract: the cross-check of the tensor-core kernel, and the path for shapes outside its
// domain (width < 8, channel counts that are not multiples of 8, fp32 storage).  One thread = one (tap, co, ci), pixels serial.
template <typename T>
__global__ void wgrad_ref_kernel(const T* __restrict__ x, const T* __restrict__ gz, float* __restrict__ dw, int n_img, int h, int w, int cin,
                                 int cout, int per_sample) {
  const long total = 9L * cout * cin;
  const int s = blockIdx.y;   // per-sample: image; shared: 0
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ci = idx % cin, co = (idx / cin) % cout, tap = idx / (static_cast<long>(cin) * cout);
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    float acc = 0.f;
    const int n0 = per_sample ? s : 0, n1 = per_sample ? s + 1 : n_img;
    for (int n = n0; n < n1; ++n)
      for (int p = 0; p < h; ++p) {
        const int ph = p + dy;
        if (ph < 0 || ph >= h) continue;
        for (int qx = 0; qx < w; ++qx) {
          const int qw = qx + dx;
          if (qw < 0 || qw >= w) continue;
          acc = fmaf(to_f32(gz[((static_cast<long>(n) * h + p) * w + qx) * cout + co]), to_f32(x[((static_cast<long>(n) * h + ph) * w + qw) * cin + ci]), acc);
        }
      }
    dw[(static_cast<long>(s) * 9 + tap) * cout * cin + static_cast<long>(co) * cin + ci] += acc;
  }
}

// db[c] += sum over n, h, w of gz (the bias gradient of conv3x3+bias+ReLU; gz is already masked by the ReLU)
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ gz, float* __restrict__ db, long pixels, int c) {
  extern __shared__ float sred[];
  const int ch = threadIdx.x % c;          // blockDim.x is a multiple of c (launcher)
  float acc = 0.f;
  const long step = static_cast<long>(gridDim.x) * (blockDim.x / c);
  for (long p = blockIdx.x * static_cast<long>(blockDim.x / c) + threadIdx.x / c; p < pixels; p += step) acc += to_f32(gz[p * c + ch]);
  sred[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < c) {
    float t = 0.f;
    for (int i = threadIdx.x; i < blockDim.x; i += c) t += sred[i];
    atomicAdd(db + threadIdx.x, t);
  }
}

// First conv of VGG / encoder (3 input channels, fp32 NCHW input): dw[co][ci][ky][kx] += sum gz[n][h][w][co] * x[n][ci][h+ky-1][w+kx-1].
// One thread = one output channel, 27 accumulators in registers; a block walks a strip of pixels (the 27 input taps of a pixel are
// the same for every thread: broadcast loads), then one atomicAdd per (thread, tap).
template <typename T>
__global__ void conv_c3_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ gz, float* __restrict__ dw, int n_img, int h, int w, int cout) {
  const long pixels = static_cast<long>(n_img) * h * w;
  const long per_block = (pixels + gridDim.x - 1) / gridDim.x;
  const long p0 = blockIdx.x * per_block, p1 = min(pixels, p0 + per_block);
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.f;
    for (long p = p0; p < p1; ++p) {
      const int n = p / (static_cast<long>(h) * w), r = p % (static_cast<long>(h) * w), py = r / w, px = r % w;
      const float g = to_f32(gz[p * cout + co]);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int yy = py + ky - 1, xx = px + kx - 1;
            const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(x + ((static_cast<long>(n) * 3 + ci) * h + yy) * w + xx) : 0.f;
            acc[(ci * 3 + ky) * 3 + kx] = fmaf(g, v, acc[(ci * 3 + ky) * 3 + kx]);
          }
    }
#pragma unroll
    for (int i = 0; i < 27; ++i) atomicAdd(dw + static_cast<long>(co) * 27 + i, acc[i]);
  }
}

// ModulatedConv2d: from the per-sample GEMM result G[n][tap][co][ci] = sum_p gz[n][p][co] x[n][p+tap][ci]  (gz = d * dL/dy, x the
// UNmodulated input) to the gradient of the shared weight (oracle/stylegan2.py modulated_conv2d):
//   w'[n] = Wb * s[n],  d[n][co] = rsqrt(sum w'^2 + eps),  y = d * conv(w', x)         (Wb = scale * W)
//   dL/dw'[n][co][ci][tap] = s-free direct term G' - gdacc[n][co] * d[n][co]^2 * w'[n][co][ci][tap]
//   dL/dWb[tap][co][ci]    = sum_n s[n][ci] * (G[n][tap][co][ci] - gdacc[n][co] * d[n][co]^2 * Wb[tap][co][ci] * s[n][ci])
// (G computed on the unmodulated x carries one factor s[n][ci] less than dL/dw', hence the single s in front.)
__global__ void modconv_wgrad_finish_kernel(const float* __restrict__ G, const float* __restrict__ wb, const float* __restrict__ s, int s_stride,
                                            const float* __restrict__ d, const float* __restrict__ gdacc, float* __restrict__ dwb, int n_img,
                                            int cout, int cin, int demodulate) {
  const long total = 9L * cout * cin;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ci = idx % cin, co = (idx / cin) % cout;
    const float w = wb[idx];
    float acc = 0.f;
    for (int n = 0; n < n_img; ++n) {
      const float sv = s[static_cast<long>(n) * s_stride + ci];
      float t = G[static_cast<long>(n) * total + idx];
      if (demodulate) {
        const float dv = d[static_cast<long>(n) * cout + co];
        t -= gdacc[static_cast<long>(n) * cout + co] * dv * dv * w * sv;
      }
      acc = fmaf(sv, t, acc);
    }
    dwb[idx] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_nhwc(EncodeTiledFn enc, CUtensorMap* map, const void* base, int n, int h, int w, int c, int TW, int rows) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)rows, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace

extern "C" int sfk_conv3x3_wgrad(const void* x, const void* gz, float* dw, int n, int h, int w, int cin, int cout, int per_sample, int use_ref,
                                 int* err, sfk_stream_t stream) {
  SFK_REQUIRE(x && gz && dw && n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, SFK_E_ARG, "conv3x3_wgrad: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = per_sample ? n : 1;
  const bool tc_ok = !sfk_act_f32() && !use_ref && w >= 8 && h >= 2 && cin % 8 == 0 && cout % 8 == 0 && sfk_aligned16(x) && sfk_aligned16(gz);
  if (!tc_ok) {
    const long total = 9L * cout * cin;
    long blocks = (total + 127) / 128;
    if (blocks > 4096) blocks = 4096;
    if (sfk_act_f32())
      wgrad_ref_kernel<float><<<dim3(static_cast<unsigned>(blocks), S), 128, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(gz), dw, n, h, w, cin, cout, per_sample);
    else
      wgrad_ref_kernel<__nv_bfloat16><<<dim3(static_cast<unsigned>(blocks), S), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gz), dw, n, h, w, cin, cout, per_sample);
    return sfk_check_launch("wgrad_ref_kernel");
  }
  EncodeTiledFn enc = encode_fn();
  SFK_REQUIRE(enc != nullptr, SFK_E_DRIVER, "conv3x3_wgrad: cuTensorMapEncodeTiled unavailable");
  WgradArgs k;
  memset(&k, 0, sizeof(k));
  k.dw = dw; k.n_img = n; k.h = h; k.w = w; k.cout = cout; k.cin = cin; k.per_sample = per_sample ? 1 : 0; k.err = err;
  k.TW = w > 8 ? 16 : 8;
  k.TH = kTilePx / k.TW;
  k.tiles_w = (w + k.TW - 1) / k.TW;
  k.tiles_h = (h + k.TH - 1) / k.TH;
  k.m_boxes = cout > 64 ? 2 : 1;
  k.n_boxes = cin > 64 ? 2 : 1;
  const int co_blocks = (cout + 127) / 128;
  k.ci_blocks = (cin + 64 * k.n_boxes - 1) / (64 * k.n_boxes);
  k.a_box_bytes = kTilePx * 128;
  k.b_box_bytes = (k.TH + 2) * k.TW * 128;
  k.stage_bytes = k.m_boxes * k.a_box_bytes + k.n_boxes * k.b_box_bytes;   // every box is a multiple of 1024 bytes
  SFK_REQUIRE(encode_nhwc(enc, &k.mapA, gz, n, h, w, cout, k.TW, k.TH) == 0 && encode_nhwc(enc, &k.mapB, x, n, h, w, cin, k.TW, k.TH + 2) == 0,
              SFK_E_DRIVER, "conv3x3_wgrad: cuTensorMapEncodeTiled failed");
  // splits: fill the machine (one CTA per SM: 512 TMEM columns), at least 2 tiles per split where there are that many
  const int items = co_blocks * k.ci_blocks * 3;
  const int tiles = k.tiles_h * k.tiles_w * (per_sample ? 1 : n);
  int spl = (sfk_num_sms() + items * S - 1) / (items * S);
  if (spl > (tiles + 1) / 2) spl = (tiles + 1) / 2;
  if (spl < 1) spl = 1;
  k.splits = spl * S;
  int stages = (200 * 1024) / k.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > (tiles + spl - 1) / spl) stages = (tiles + spl - 1) / spl;
  if (stages < 1) stages = 1;
  k.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * k.stage_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  if (e != cudaSuccess) return static_cast<int>(e);
  wgrad_tc_kernel<<<dim3(static_cast<unsigned>(k.splits), static_cast<unsigned>(items)), kThreads, smem, st>>>(k);
  return sfk_check_launch("wgrad_tc_kernel");
}

extern "C" int sfk_bias_grad(const void* gz, float* db, int n, int hw, int c, sfk_stream_t stream) {
  SFK_REQUIRE(gz && db && n > 0 && hw > 0 && c > 0 && c <= 1024, SFK_E_ARG, "bias_grad: bad args");
  int threads = c;
  while (threads * 2 <= 512) threads *= 2;
  threads = (threads / c) * c;
  const long pixels = static_cast<long>(n) * hw;
  long blocks = (pixels + (threads / c) * 8 - 1) / ((threads / c) * 8);
  if (blocks > 2 * sfk_num_sms()) blocks = 2 * sfk_num_sms();
  if (blocks < 1) blocks = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (sfk_act_f32())
    bias_grad_kernel<float><<<static_cast<unsigned>(blocks), threads, threads * sizeof(float), st>>>(static_cast<const float*>(gz), db, pixels, c);
  else
    bias_grad_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), threads, threads * sizeof(float), st>>>(static_cast<const __nv_bfloat16*>(gz), db, pixels, c);
  return sfk_check_launch("bias_grad_kernel");
}

extern "C" int sfk_modconv_wgrad_finish(const float* G, const float* wb, const float* s, int s_stride, const float* d, const float* gdacc,
                                        float* dwb, int n, int cout, int cin, int demodulate, sfk_stream_t stream) {
  SFK_REQUIRE(G && wb && s && dwb && (!demodulate || (d && gdacc)) && n > 0 && cout > 0 && cin > 0, SFK_E_ARG, "modconv_wgrad_finish: bad args");
  const long total = 9L * cout * cin;
  long blocks = (total + 255) / 256;
  if (blocks > 8 * sfk_num_sms()) blocks = 8 * sfk_num_sms();
  modconv_wgrad_finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(G, wb, s, s_stride, d, gdacc, dwb, n, cout, cin,
                                                                                                          demodulate);
  return sfk_check_launch("modconv_wgrad_finish_kernel");
}

extern "C" int sfk_conv_c3_wgrad(const float* x, const void* gz, float* dw, int n, int h, int w, int cout, sfk_stream_t stream) {
  SFK_REQUIRE(x && gz && dw && n > 0 && h > 0 && w > 0 && cout > 0, SFK_E_ARG, "conv_c3_wgrad: bad args");
  const long pixels = static_cast<long>(n) * h * w;
  long blocks = (pixels + 255) / 256;
  if (blocks > 8L * sfk_num_sms()) blocks = 8L * sfk_num_sms();
  const int threads = cout >= 128 ? 128 : (cout >= 64 ? 64 : 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (sfk_act_f32())
    conv_c3_wgrad_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, static_cast<const float*>(gz), dw, n, h, w, cout);
  else
    conv_c3_wgrad_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, static_cast<const __nv_bfloat16*>(gz), dw, n, h, w, cout);
  return sfk_check_launch("conv_c3_wgrad_kernel");
}
