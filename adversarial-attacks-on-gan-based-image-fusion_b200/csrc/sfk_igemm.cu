// Implicit-GEMM convolution for sm_100a: TMA (cp.async.bulk.tensor) -> 128B/64B/32B-swizzled smem ->
// tcgen05.mma (one issuing thread, fp32 accumulators in TMEM) -> tcgen05.ld epilogue.
//
// GEMM view (see include/sfk.h): M = 128 output positions (a TH x TW patch of one image),
// N = block_n output channels per accumulator, K = taps x Cin walked in k-steps of KC channels.
// Because the activation tensor is NHWC, the box {KC, TW, TH, 1, 1} of a 5-D tensor map lands in
// shared memory as 128 rows x (KC*2) bytes -- exactly the K-major swizzled operand layout
// tcgen05.mma wants -- and the spatial shift (dy,dx) of a filter tap is just a coordinate offset;
// out-of-bounds rows are zero-filled by TMA, which implements the conv padding for free.
//
// Warp roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM alloc/dealloc,
// warps4-7 = epilogue (TMEM lanes 32*(warp%4)..).  Two CTAs fit per SM so one CTA's epilogue
// overlaps the other's main loop.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "sfk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr long long kTimeoutCycles = 400000000LL;  // ~0.2 s: a hung pipeline flags an error instead of hanging the GPU

struct KTap {
  int dy, dx, plane, acc, brow, first;
};

struct __align__(64) IgemmKArgs {
  CUtensorMap mapA;
  CUtensorMap mapB;
  int n_img, out_h, out_w, out_c;
  int TH, TW, tiles_h, tiles_w;
  int KC, num_cblk, block_n, num_acc, num_taps, stages;
  int b_per_sample;
  int a_stage_bytes, b_stage_bytes;
  int layout_type;  // UMMA LayoutType: 2 = SW128, 4 = SW64, 6 = SW32
  int sbo_bytes;
  int tmem_cols;
  int flags;
  int vec_stride;
  int out_d2s;
  float noise_w;
  __nv_bfloat16* out;
  const float* dscale;
  const float* bias;
  const float* noise;
  const __nv_bfloat16* xin;
  const float* colscale;
  float* gs;
  int* err;
  KTap taps[SFK_MAX_TAPS];
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
// Bounded wait: returns false (and raises *err) if the barrier never flips.  The abort flag lives in global memory, so it
// is polled only once per 64 failed probes (try_wait itself suspends the thread for a hardware time slice): a load in the
// spin loop would add its full latency to every wake-up.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if (mbar_try_wait(bar, parity)) return true;
    if (clock64() - t0 > kTimeoutCycles || (err && *reinterpret_cast<volatile int*>(err) != 0)) {
      if (err) atomicExch(err, 1);
      return false;
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// shared -> global tensor store of one box (bulk async group; completion tracked with cp.async.bulk.wait_group[.read])
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ uint4 shfl_xor_u4(uint4 v, int m) {
  return make_uint4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m), __shfl_xor_sync(0xffffffffu, v.z, m),
                    __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);         // start address, bits [0,14)
  d |= static_cast<uint64_t>(1u) << 16;                      // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset between 8-row groups, bits [32,46)
  d |= static_cast<uint64_t>(1u) << 46;                      // descriptor version = 1 on sm_100
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;        // swizzle mode, bits [61,64)
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
// kind::tf32: fp32 bit patterns in shared memory, read as tf32 (10-bit mantissa); one instruction covers K = 8 elements = 32 bytes,
// so the operand descriptors advance exactly like the bf16 ones (K = 16 elements = 32 bytes)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
template <typename T>
__device__ __forceinline__ void umma_t(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (std::is_same<T, float>::value) umma_tf32(tmem_c, adesc, bdesc, idesc, accum);
  else umma_bf16(tmem_c, adesc, bdesc, idesc, accum);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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

// Sum each of 16 per-lane values over the 32 lanes with 16 shuffles; afterwards lane l holds the total
// of column  8*b0 + 4*b1 + 2*b2 + b3  (b_i = bit i of l).
__device__ __forceinline__ float warp_colsum16(float* v, int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 1 << step;
    const int half = 8 >> step;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < half) {
        const float keep = upper ? v[i + half] : v[i];
        const float send = upper ? v[i] : v[i + half];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}
__device__ __forceinline__ int colsum16_column(int lane) {
  return 8 * (lane & 1) + 4 * ((lane >> 1) & 1) + 2 * ((lane >> 2) & 1) + ((lane >> 3) & 1);
}

// The epilogue arithmetic for one output position and 16 consecutive channels.  Shared by the
// tensor-core kernel and the CUDA-core cross-check kernel so both define the same operator.
template <typename Args>
__device__ __forceinline__ void epilogue16(const Args& a, int n, int acc, int h, int w, int col0, bool valid, float* v,
                                           float* gsdot /* 16 products for the style-gradient reduction */) {
  const int flags = a.flags;
  long off, vec, noise_idx = static_cast<long>(h) * a.out_w + w;
  int bias0 = col0;
  if (a.out_d2s) {   // depth-to-space: column block -> output phase
    const int Cq = a.out_c / 4, ph = col0 / Cq, ch = col0 % Cq;
    const long fh = 2L * h + (ph >> 1), fw = 2L * w + (ph & 1);
    off = ((static_cast<long>(n) * 2 * a.out_h + fh) * (2 * a.out_w) + fw) * Cq + ch;
    vec = static_cast<long>(n) * Cq + ch;
    bias0 = ch;
    noise_idx = fh * (2 * a.out_w) + fw;
  } else {
    const long pix = ((static_cast<long>(n) * a.num_acc + acc) * a.out_h + h) * a.out_w + w;
    off = pix * a.out_c + col0;
    vec = static_cast<long>(n) * a.out_c + col0;
  }
  const long svec = static_cast<long>(n) * a.vec_stride + col0;  // colscale / gs rows may live inside a wider [n][s_dim] array
  if (flags & SFK_EP_DSCALE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= __ldg(a.dscale + vec + i);
  }
  if ((flags & SFK_EP_NOISE) && valid) {
    const float nz = a.noise_w * __ldg(a.noise + noise_idx);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += nz;
  }
  if (flags & SFK_EP_BIAS) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + bias0 + i);
  }
  if (flags & SFK_EP_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (flags & SFK_EP_LRELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = lrelu_fwd(v[i]);
  }
  if (flags & SFK_EP_LRELU_RAW) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.2f * v[i]);
  }
  if (flags & (SFK_EP_XMASK | SFK_EP_GSDOT)) {
    float x[16];
    if (valid) {
      load8(a.xin + off, x);
      load8(a.xin + off + 8, x + 8);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = 0.f;
    }
    if (flags & SFK_EP_GSDOT) {
#pragma unroll
      for (int i = 0; i < 16; ++i) gsdot[i] = x[i] * v[i];
    }
    if (flags & SFK_EP_XMASK) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = x[i] > 0.f ? v[i] : 0.f;
    }
  }
  if (flags & SFK_EP_COLSCALE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= __ldg(a.colscale + svec + i);
  }
  if (valid) {
    if (flags & SFK_EP_ACCUM) {
      float o[16];
      load8p(a.out + off, o);
      load8p(a.out + off + 8, o + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += o[i];
    }
    store8(a.out + off, v);
    store8(a.out + off + 8, v + 8);
  }
}

__global__ void __launch_bounds__(kThreads, 2) igemm_tc_kernel(const __grid_constant__ IgemmKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float gs_acc[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + a.stages * a.a_stage_bytes;

  // tile coordinates
  const int bx = blockIdx.x;
  const int tw_i = bx % a.tiles_w;
  const int th_i = (bx / a.tiles_w) % a.tiles_h;
  const int n = bx / (a.tiles_w * a.tiles_h);
  const int h0 = th_i * a.TH, w0 = tw_i * a.TW;
  const int n0 = blockIdx.y * a.block_n;
  const int num_k = a.num_cblk * a.num_taps;

  if (threadIdx.x < 256) gs_acc[threadIdx.x] = 0.f;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = static_cast<uint32_t>(128 * a.KC * 2 + a.block_n * a.KC * 2);
      int ks = 0;
      bool ok = true;
      for (int cb = 0; cb < a.num_cblk && ok; ++cb) {
        for (int t = 0; t < a.num_taps; ++t, ++ks) {
          const int stage = ks % a.stages;
          const uint32_t phase = (ks / a.stages) & 1;
          if (!mbar_wait(&empty_bar[stage], phase ^ 1, a.err)) {
            ok = false;
            break;
          }
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          const KTap& tp = a.taps[t];
          tma_load_5d(smem_a + stage * a.a_stage_bytes, &a.mapA, &full_bar[stage], cb * a.KC, w0 + tp.dx, h0 + tp.dy,
                      tp.plane, n);
          tma_load_3d(smem_b + stage * a.b_stage_bytes, &a.mapB, &full_bar[stage], cb * a.KC, tp.brow + n0,
                      a.b_per_sample ? n : 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32 (bit4), A=B=bf16 (bits 7,10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a.block_n >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);
      const int kslices = a.KC / 16;
      int ks = 0;
      bool ok = true;
      for (int cb = 0; cb < a.num_cblk && ok; ++cb) {
        for (int t = 0; t < a.num_taps; ++t, ++ks) {
          const int stage = ks % a.stages;
          const uint32_t phase = (ks / a.stages) & 1;
          if (!mbar_wait(&full_bar[stage], phase, a.err)) {
            ok = false;
            break;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const KTap& tp = a.taps[t];
          const uint64_t adesc = make_smem_desc(smem_a + stage * a.a_stage_bytes, a.sbo_bytes, a.layout_type);
          const uint64_t bdesc = make_smem_desc(smem_b + stage * a.b_stage_bytes, a.sbo_bytes, a.layout_type);
          const uint32_t tmem_c = tmem_base + static_cast<uint32_t>(tp.acc * a.block_n);
          for (int k = 0; k < kslices; ++k) {
            const uint32_t accum = (cb == 0 && tp.first && k == 0) ? 0u : 1u;
            // advancing K by 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr>>4) field
            umma_bf16(tmem_c, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, accum);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        }
      }
      umma_commit(&tmem_full_bar);  // accumulators complete
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const bool ok = mbar_wait(&tmem_full_bar, 0, a.err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int th = row / a.TW, tw = row % a.TW;
    const int h = h0 + th, w = w0 + tw;
    const bool valid = ok && (h < a.out_h) && (w < a.out_w);
    const int chunks = a.block_n / 16;
    const int mycol = colsum16_column(lane);
    for (int acc = 0; acc < a.num_acc; ++acc) {
      for (int c = 0; c < chunks; ++c) {
        float v[16], gsd[16];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * a.block_n + c * 16);
        tmem_ld16(taddr, v);
        epilogue16(a, n, acc, h, w, n0 + c * 16, valid, v, gsd);
        if (a.flags & SFK_EP_GSDOT) {
          const float tot = warp_colsum16(gsd, lane);
          if (lane < 16) atomicAdd(&gs_acc[c * 16 + mycol], tot);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if ((a.flags & SFK_EP_GSDOT) && threadIdx.x < a.block_n) {
    atomicAdd(a.gs + static_cast<long>(n) * a.vec_stride + n0 + threadIdx.x, gs_acc[threadIdx.x]);
  }
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
  }
}

// =============================================================================================
// Persistent variant (the default): each CTA owns one (image, N-block) and walks its output tiles.
//   * the TMA/MMA pipeline runs across tile boundaries and the accumulator is double-buffered in TMEM,
//     so the epilogue of tile i overlaps the main loop of tile i+1;
//   * filter taps that differ only in dy share ONE activation load: the box is TH+span rows tall and a
//     tap's operand is the same smem tile entered (dy-dy_min)*TW rows further down (TW is a multiple of
//     8 rows, so the swizzle phase is preserved) -- 3 loads instead of 9 for a 3x3 conv;
//   * small weight sets (<= 72 KB for this CTA's N-block) are loaded once and stay resident in smem;
//   * per-column epilogue vectors are staged in smem once per CTA; the style-gradient partial sums
//     are accumulated in smem over all tiles and flushed with block_n atomics per CTA.
constexpr int kMaxGroups = 9;
constexpr int kMaxGroupTaps = 9;

struct KGroup {
  int plane, dx_min, dy_min, map, bytes, ntaps;   // map: which activation tensor map (box height); bytes: box size
  int roff[kMaxGroupTaps];                        // operand start, in smem rows, of each tap inside the loaded box
  int acc[kMaxGroupTaps], brow[kMaxGroupTaps], first[kMaxGroupTaps], bidx[kMaxGroupTaps];
  // host-precomputed issue constants, read straight into uniform registers by the MMA warp
  int a16[kMaxGroupTaps], b16[kMaxGroupTaps], col[kMaxGroupTaps];
};

// what the kernel reads of a load group (the argument block must stay below 4 KB: beyond that the parameters leave the fast
// constant bank -- measured: every launch 10-25 % slower with a 4156-byte block)
struct KGroupDev {
  int plane, dx_min, dy_min, map, bytes, ntaps;
  int brow[kMaxGroupTaps], bidx[kMaxGroupTaps];
};

struct __align__(64) Igemm2Args {
  CUtensorMap mapA[4];  // box height TH + 0..3 rows
  CUtensorMap mapB;
  CUtensorMap mapO;     // output, for the staged TMA-store epilogue (ts != 0): box {ts_slabw ch, TW, 1, 32/TWB rows, 1}
  int xs;   // quad-transposed direct stores (see the epilogue)
  int m2;   // two M tiles (16 tile rows) per pipeline stage: every weight tile loaded into shared memory feeds 256 output rows
  int ts, ts_slabw, ts_nbuf, ts_off;   // staging: per epilogue warp ts_nbuf buffers of 32 rows x ts_slabw bf16 at smem_base + ts_off
  int n_img, out_h, out_w, out_c;
  int TH, TW, TWB, tiles_h, tiles_w, n_blocks;   // TWB = box width = row pitch of the M index; TW <= TWB useful columns
  int KC, num_cblk, block_n, num_acc, num_taps, num_groups, stages, acc_stages;
  int b_per_sample, b_resident, dual_issue;
  int ksplit;              // fp32 storage: k-blocks rotate over `ksplit` partial accumulators that the epilogue sums (see the MMA warp)
  int passes, b_samples;   // passes == 3: split-tf32 (A.hi*B.hi + A.lo*B.hi + A.hi*B.lo); the lo halves sit n_img images / b_samples weight sets further on
  int out_d2s, a_s2d, cpa, cq_log2;   // fused resampling (see sfk.h); cpa = k-blocks per row phase of the space-to-depth input
  int a_stage_bytes, b_tap_bytes, b_stage_bytes, row_bytes;
  int layout_type, sbo_bytes, tmem_cols, flags, vec_stride;
  float noise_w;
  void* out;          // bf16 or fp32 (kernel template parameter)
  const float* dscale;
  const float* bias;
  const float* noise;
  const void* xin;
  const float* colscale;
  float* gs;
  int* err;
  KGroupDev groups[kMaxGroups];
  // taps flattened in issue order for the MMA warp (kept in registers): operand offsets (16 B units), TMEM column,
  // flags bit0 = first MMA into its accumulator, bit1 = first tap of a load group (wait for data), bit2 = last (release slot)
  int f_a16[SFK_MAX_TAPS], f_b16[SFK_MAX_TAPS], f_col[SFK_MAX_TAPS], f_flags[SFK_MAX_TAPS];
  // Merged taps (build_plan): taps of one load group that read the SAME activation operand and whose accumulators sit in adjacent
  // TMEM column blocks are issued as ONE tcgen05.mma of f_wm * block_n columns over their adjacent weight tiles (f_wm = 0: folded
  // into the tap before it).  acc_blk maps an accumulator (= output plane) to its TMEM column block.
  int f_wm[SFK_MAX_TAPS];
  int acc_blk[4];
  int merged;   // any f_wm != 1
  int early;    // the epilogue releases the accumulator stage as soon as its last chunk sits in registers (before the arithmetic and stores)
};

// role-level cycle accounting (only when SFK_EP_PROFILE is set): [0] producer waiting for a free slot, [1] producer total,
// [2] MMA waiting for data, [3] MMA waiting for a free accumulator, [4] MMA total, [5] epilogue waiting for the accumulator,
// [6] epilogue total, [7] tiles
__device__ unsigned long long g_role_cycles[8];

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
#define SFK_EP_PROFILE (1 << 16)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// FCT >= 0: the epilogue flag set is a compile-time constant (the six combinations the engines use are instantiated, so untaken
// epilogue paths and their index arithmetic disappear); FCT < 0: flags are read at run time.
// T: storage type of activations and weights.  __nv_bfloat16 -> kind::f16 MMAs (the product path); float -> kind::tf32 MMAs on fp32
// storage (the parity mode), optionally as three passes over hi/lo-split operands for fp32-class products.
// VAR >= 0: the launch variant is a compile-time constant too (bit 0 depth-to-space output, bit 1 two M tiles per stage, bit 2
// quad-transposed stores; never the staged TMA store), so that the per-tile code of the hot launches carries none of the other
// variants' branches and address arithmetic (source-level profile, round 1: ~640 instructions per 32-column tile, 160 of them
// arithmetic); VAR < 0: read from the arguments at run time.
constexpr int kVarD2S = 1, kVarM2 = 2, kVarXS = 4, kVarMG = 8;   // kVarMG: merged taps (per-tap MMA widths, permuted accumulator blocks)
template <int FCT, typename T, int VAR>
__global__ void __launch_bounds__(kThreads, 2) igemm_tc2_kernel(const __grid_constant__ Igemm2Args a) {
  constexpr bool kF32 = std::is_same<T, float>::value;
  const bool v_d2s = VAR >= 0 ? (VAR & kVarD2S) != 0 : a.out_d2s != 0;
  const bool v_m2 = VAR >= 0 ? (VAR & kVarM2) != 0 : a.m2 != 0;
  const bool v_xs = VAR >= 0 ? (VAR & kVarXS) != 0 : a.xs != 0;
  const bool v_ts = VAR >= 0 ? false : a.ts != 0;
  const bool v_mg = VAR >= 0 ? (VAR & kVarMG) != 0 : a.merged != 0;
  const int m2n = v_m2 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint64_t bres_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float gs_acc[256];
  __shared__ __align__(16) float col_dscale[256];
  __shared__ __align__(16) float col_bias[256];
  __shared__ __align__(16) float col_scale[256];
  // ready-made 64-bit operand descriptors, so that issuing a tap is: 2 x LDS.64 -> uniform registers -> tcgen05.mma
  //   s_adesc[stage][tap]                      s_bdesc[resident ? channel block : stage][tap]
  __shared__ uint64_t s_adesc[kMaxStages * SFK_MAX_TAPS];
  __shared__ uint64_t s_bdesc[kMaxStages * SFK_MAX_TAPS];
  __shared__ int s_colf[SFK_MAX_TAPS];   // (TMEM column << 1) | first-MMA-into-accumulator
  __shared__ int s_wm[SFK_MAX_TAPS];     // merged taps: width multiplier (0 = issued as part of the tap before it)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + a.stages * a.a_stage_bytes;  // per-stage B tiles, or the resident weight set

  const int grp = blockIdx.y;  // (image, N-block)
  const int n = grp / a.n_blocks;
  const int n0 = (grp % a.n_blocks) * a.block_n;
  const int tiles_per_group = a.tiles_h * a.tiles_w;
  const int bs = a.b_per_sample ? n : 0;

  gs_acc[threadIdx.x] = 0.f;
  if (threadIdx.x < kMaxStages * SFK_MAX_TAPS) {
    const int slot = threadIdx.x / SFK_MAX_TAPS, t = threadIdx.x % SFK_MAX_TAPS;
    const uint32_t base_a = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t base_b = base_a + a.stages * a.a_stage_bytes;
    const uint64_t hi = make_smem_desc(0, a.sbo_bytes, a.layout_type);
    uint64_t da = 0, db = 0;
    if (t < a.num_taps) {
      if (slot < a.stages) da = hi | static_cast<uint64_t>(((base_a + slot * a.a_stage_bytes) >> 4) + static_cast<uint32_t>(a.f_a16[t]));
      if (a.b_resident) {   // slot = channel block (+ num_cblk for the lo halves of the split-tf32 passes)
        if (slot < a.num_cblk * (a.passes == 3 ? 2 : 1)) db = hi | static_cast<uint64_t>(((base_b + slot * a.num_taps * a.b_tap_bytes) >> 4) + static_cast<uint32_t>(a.f_b16[t]));
      } else if (slot < a.stages) {
        db = hi | static_cast<uint64_t>(((base_b + slot * a.b_stage_bytes) >> 4) + static_cast<uint32_t>(a.f_b16[t]));
      }
    }
    s_adesc[threadIdx.x] = da;
    s_bdesc[threadIdx.x] = db;
    if (slot == 0) s_colf[t] = t < a.num_taps ? ((a.f_col[t] << 1) | (a.f_flags[t] & 1)) : 0;
    if (slot == 1) s_wm[t] = t < a.num_taps ? a.f_wm[t] : 0;
  }
  if (threadIdx.x < a.block_n) {
    const int c = n0 + threadIdx.x;
    const int cw = v_d2s ? a.out_c / 4 : a.out_c;   // depth-to-space: the 4 phases share the per-channel vectors
    col_dscale[threadIdx.x] = (a.flags & SFK_EP_DSCALE) ? a.dscale[static_cast<long>(n) * cw + c % cw] : 1.f;
    col_bias[threadIdx.x] = (a.flags & SFK_EP_BIAS) ? a.bias[c % cw] : 0.f;
    col_scale[threadIdx.x] = (a.flags & SFK_EP_COLSCALE) ? a.colscale[static_cast<long>(n) * a.vec_stride + c] : 1.f;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapA[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapA[a.groups[0].map])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);   // one arrival per epilogue warp
    }
    mbar_init(&bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      bool ok = true;
      if (a.b_resident) {
        const int halves = a.passes == 3 ? 2 : 1;
        mbar_expect_tx(&bres_bar, static_cast<uint32_t>(halves * a.num_cblk * a.num_taps * a.block_n * a.row_bytes));
        for (int hf = 0; hf < halves; ++hf)
          for (int cb = 0; cb < a.num_cblk; ++cb)
            for (int g = 0; g < a.num_groups; ++g)
              for (int j = 0; j < a.groups[g].ntaps; ++j)
                tma_load_3d(smem_b + ((hf * a.num_cblk + cb) * a.num_taps + a.groups[g].bidx[j]) * a.b_tap_bytes, &a.mapB, &bres_bar,
                            cb * a.KC, a.groups[g].brow[j] + n0, bs + hf * a.b_samples);
      }
      int stage = 0;           // ring slot and its phase advance incrementally (no division by the run-time stage count per k-step)
      uint32_t phase = 0;
      const bool prof = (a.flags & SFK_EP_PROFILE) != 0;
      long long t_wait = 0;
      const long long t_start = clock64();
      // per-group constants live in registers (fully unrolled group loop); tile coordinates advance without divisions
      int Gdx[kMaxGroups], Gdy[kMaxGroups], Gpl[kMaxGroups], Gby[kMaxGroups], Gnt[kMaxGroups];
      uint64_t Gmap[kMaxGroups];
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        Gdx[g] = a.groups[g].dx_min;
        Gdy[g] = a.groups[g].dy_min;
        Gpl[g] = a.groups[g].plane;
        Gnt[g] = a.groups[g].ntaps;
        Gby[g] = a.groups[g].bytes + (a.b_resident ? 0 : a.groups[g].ntaps * a.block_n * a.row_bytes);
        Gmap[g] = reinterpret_cast<uint64_t>(&a.mapA[a.groups[g].map]);
      }
      const int ngroups = a.num_groups;
      int t_h = static_cast<int>(blockIdx.x) / a.tiles_w, t_w = static_cast<int>(blockIdx.x) % a.tiles_w;
      for (int tile = blockIdx.x; tile < tiles_per_group && ok; tile += gridDim.x) {
        const int h0 = t_h * a.TH, w0 = t_w * a.TW;
        t_w += gridDim.x;
        while (t_w >= a.tiles_w) {
          t_w -= a.tiles_w;
          ++t_h;
        }
        for (int cbx = 0; cbx < a.num_cblk * a.passes && ok; ++cbx) {
          // split-tf32: pass 0 = A.hi x B.hi, pass 1 = A.lo x B.hi, pass 2 = A.hi x B.lo
          const int pass = cbx / a.num_cblk, cb = cbx - pass * a.num_cblk;
          const int na = n + (pass == 1 ? a.n_img : 0), nb = bs + (pass == 2 ? a.b_samples : 0);
#pragma unroll
          for (int g = 0; g < kMaxGroups; ++g) {
            if (g < ngroups && ok) {
              const long long tw0 = prof ? clock64() : 0;
              ok = mbar_wait(&empty_bar[stage], phase ^ 1, a.err);
              if (prof) t_wait += clock64() - tw0;
              if (ok) {
                mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(Gby[g]));
                if (a.a_s2d)   // dims {2Cq, W, row phase, H, N}: k-block cb = (row phase, part of the pixel pair)
                  tma_load_5d(smem_a + stage * a.a_stage_bytes, reinterpret_cast<const CUtensorMap*>(Gmap[g]), &full_bar[stage],
                              (cb % a.cpa) * a.KC, w0 + Gdx[g], cb / a.cpa, h0 + Gdy[g], na);
                else
                  tma_load_5d(smem_a + stage * a.a_stage_bytes, reinterpret_cast<const CUtensorMap*>(Gmap[g]), &full_bar[stage], cb * a.KC,
                              w0 + Gdx[g], h0 + Gdy[g], Gpl[g], na);
                if (!a.b_resident) {
                  for (int j = 0; j < Gnt[g]; ++j)
                    tma_load_3d(smem_b + stage * a.b_stage_bytes + j * a.b_tap_bytes, &a.mapB, &full_bar[stage], cb * a.KC,
                                a.groups[g].brow[j] + n0, nb);
                }
              }
              if (++stage == a.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
      if (prof) {
        atomicAdd(&g_role_cycles[0], static_cast<unsigned long long>(t_wait));
        atomicAdd(&g_role_cycles[1], static_cast<unsigned long long>(clock64() - t_start));
      }
    }
  } else if (warp == 1 || (warp == 3 && a.dual_issue)) {
    // ===================== MMA issuer(s) =====================
    // Issue is instruction-latency bound for small N, so with a double-buffered accumulator TWO warps issue: warp 1 owns the
    // even tiles (accumulator stage 0), warp 3 the odd ones (stage 1).  MMAs into different accumulators are independent,
    // each warp consumes exactly the ring slots of its own tiles, and every barrier still sees one arrival per use.
    // Only enabled when the ring holds two whole tiles (stages >= 2 * slots per tile): an mbarrier parity wait cannot tell a
    // phase from the one two wraps earlier, so the two consumers must never be more than one wrap apart.
    // The whole warp walks the loop convergently (so descriptors live in uniform registers without vote loops);
    // one elected lane issues tcgen05.mma / tcgen05.commit.
    {
      const bool leader = elect_one();
      // instruction descriptor: D = f32 (bit 4); A/B format at bits [7,10)/[10,13): 1 = bf16, 2 = tf32; both K-major
      const uint32_t fmt = kF32 ? 2u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(a.block_n >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);
      const int kslices = a.row_bytes / 32;   // one MMA consumes 32 bytes of K per row (16 bf16 / 8 tf32)
      const bool prof = (a.flags & SFK_EP_PROFILE) != 0;
      long long t_wd = 0, t_wa = 0;
      const long long t_start = clock64();
      const uint64_t desc_hi = make_smem_desc(0, a.sbo_bytes, a.layout_type);   // everything but the start address
      bool ok = true;
      if (a.b_resident) ok = mbar_wait(&bres_bar, 0, a.err);
      const int par = (warp == 3) ? 1 : 0;
      const int step = a.dual_issue ? 2 : 1;
      const int kpt = a.num_cblk * a.passes * a.num_groups;   // ring slots per tile
      const uint64_t m2_a16 = static_cast<uint64_t>((8 * a.TWB * a.row_bytes) >> 4);
      const uint32_t m2_col = static_cast<uint32_t>(a.num_acc * a.block_n);
      // ring slot / phase of this warp's next k-step, advanced incrementally (the other issuer's tiles are skipped kpt slots at a time)
      int stage = (par * kpt) % a.stages;
      uint32_t phase = static_cast<uint32_t>((par * kpt) / a.stages) & 1u;
      const int skip = (step - 1) * kpt;
      for (int it = par; blockIdx.x + it * static_cast<int>(gridDim.x) < tiles_per_group && ok; it += step) {
        const int as = a.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t aph = static_cast<uint32_t>(a.acc_stages == 2 ? (it >> 1) : it) & 1u;
        const long long ta0 = prof ? clock64() : 0;
        if (!mbar_wait(&tmem_empty_bar[as], aph ^ 1, a.err)) break;
        if (prof) t_wa += clock64() - ta0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // The tensor core adds into its fp32 accumulator with truncation, not round-to-nearest (measured: the error of a split-tf32
        // conv grows linearly with K, -1.2e-8 relative per accumulated element, tools/diag_tf32_accuracy.py), which would dominate
        // the parity mode's error.  Under fp32 storage the k-blocks therefore rotate over `ksplit` partial accumulators, each seeing
        // 1/ksplit of the additions; the epilogue sums them with ordinary fp32 adds.  bf16 storage: one accumulator (ksplit = 1).
        const int ksplit = kF32 ? a.ksplit : 1;
        const uint32_t tile_cols = static_cast<uint32_t>(m2n * a.num_acc * a.block_n);
        const uint32_t tmem_tile = tmem_base + static_cast<uint32_t>(as * ksplit) * tile_cols;
        for (int cbx = 0; cbx < a.num_cblk * a.passes && ok; ++cbx) {
          const int cb = (a.passes == 3 && cbx >= 2 * a.num_cblk) ? cbx - a.num_cblk : (cbx % a.num_cblk);   // resident B slot: lo halves follow the hi ones
          const uint32_t part = kF32 ? static_cast<uint32_t>(cbx % ksplit) * tile_cols : 0u;
          int t0 = 0;
          for (int g = 0; g < a.num_groups && ok; ++g) {
            const int nt = a.groups[g].ntaps;
            // gather every descriptor of this load group BEFORE the first MMA: operands of an in-flight tcgen05.mma stay
            // pinned in their (uniform) registers, so descriptors formed one tap at a time would serialise issue and execution
            uint64_t AD[kMaxGroupTaps], BD[kMaxGroupTaps];
            uint32_t TC[kMaxGroupTaps], AF[kMaxGroupTaps], ID[kMaxGroupTaps];
            const int a_row = stage * SFK_MAX_TAPS + t0;
            const int b_row = (a.b_resident ? cb : stage) * SFK_MAX_TAPS + t0;
#pragma unroll
            for (int j = 0; j < kMaxGroupTaps; ++j) {
              if (j < nt) {
                AD[j] = s_adesc[a_row + j];
                BD[j] = s_bdesc[b_row + j];
                TC[j] = tmem_tile + part + static_cast<uint32_t>(s_colf[t0 + j] >> 1);
                AF[j] = (cbx < ksplit && (s_colf[t0 + j] & 1)) ? 0u : 1u;   // first k-block of each partial overwrites
                if (v_mg) {   // (only the merged variant keeps per-tap instruction descriptors live: uniform registers are scarce here)
                  const uint32_t wm = static_cast<uint32_t>(s_wm[t0 + j]);
                  ID[j] = wm ? idesc + ((wm - 1u) * static_cast<uint32_t>(a.block_n >> 3) << 17) : 0u;
                }
              }
            }
            const long long td0 = prof ? clock64() : 0;
            ok = mbar_wait(&full_bar[stage], phase, a.err);
            if (prof) t_wd += clock64() - td0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader && ok) {
#pragma unroll
              for (int j = 0; j < kMaxGroupTaps; ++j) {
                if (j < nt && (!v_mg || ID[j] != 0u)) {
                  const uint32_t idj = v_mg ? ID[j] : idesc;
                  umma_t<T>(TC[j], AD[j], BD[j], idj, AF[j]);
#pragma unroll
                  for (int k = 1; k < 4; ++k)
                    if (k < kslices) umma_t<T>(TC[j], AD[j] + static_cast<uint64_t>(2 * k), BD[j] + static_cast<uint64_t>(2 * k), idj, 1u);
                  if (v_m2) {   // second M tile: rows 8..15 of the box (8 * TWB smem rows further down), its own accumulator
                    const uint64_t ad2 = AD[j] + m2_a16;
                    const uint32_t tc2 = TC[j] + m2_col;
                    umma_t<T>(tc2, ad2, BD[j], idj, AF[j]);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                      if (k < kslices) umma_t<T>(tc2, ad2 + static_cast<uint64_t>(2 * k), BD[j] + static_cast<uint64_t>(2 * k), idj, 1u);
                  }
                }
              }
            }
            __syncwarp();
            if (leader) umma_commit(&empty_bar[stage]);
            t0 += nt;
            if (++stage == a.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        __syncwarp();
        if (leader) umma_commit(&tmem_full_bar[as]);
        stage += skip;
        while (stage >= a.stages) {
          stage -= a.stages;
          phase ^= 1u;
        }
      }
      if (prof && leader) {
        atomicAdd(&g_role_cycles[2], static_cast<unsigned long long>(t_wd));
        atomicAdd(&g_role_cycles[3], static_cast<unsigned long long>(t_wa));
        atomicAdd(&g_role_cycles[4], static_cast<unsigned long long>(clock64() - t_start));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int th = row / a.TWB, tw = row % a.TWB;
    const int mycol = colsum16_column(lane);
    const int flags = FCT >= 0 ? FCT : a.flags;
    T* const outp = static_cast<T*>(a.out);
    const T* const xinp = static_cast<const T*>(a.xin);
    const bool prof = (a.flags & SFK_EP_PROFILE) != 0 && threadIdx.x == 128;
    long long t_we = 0;
    const long long t_start = clock64();
    // tile coordinates advance incrementally (no division per tile); the per-pixel noise value of the NEXT tile is fetched
    // before blocking on the accumulator of the current one, so its latency is off the critical path
    // style-gradient partial sums: for narrow accumulators (block_n <= 32) every thread keeps its row's x*gx~ products in
    // registers across ALL tiles of the CTA and the cross-row reduction happens once at the end; wider ones use the per-tile
    // butterfly (more registers would spill at 2 CTAs per SM)
    constexpr int kRegGs = 32;
    const bool reg_gs = (flags & SFK_EP_GSDOT) && a.block_n <= kRegGs;
    float gsr[kRegGs];
#pragma unroll
    for (int i = 0; i < kRegGs; ++i) gsr[i] = 0.f;
    int tile = blockIdx.x;
    int t_h = tile / a.tiles_w, t_w = tile % a.tiles_w;
    const uint32_t ts_base = smem_base + static_cast<uint32_t>(a.ts_off + q * a.ts_nbuf * 32 * a.ts_slabw * 2);
    int ts_buf = 0;
    // staged store: this warp's 32 rows are (32/TWB) tile rows of TW useful pixels, stored as one dense, swizzled box
    const int lrow = lane / a.TWB;
    const uint32_t ts_row = static_cast<uint32_t>(lrow * a.TW + tw) * static_cast<uint32_t>(a.ts_slabw * 2);
    const uint32_t ts_swz = a.ts_slabw == 64 ? ((ts_row >> 7) & 7u) : ((ts_row >> 7) & 3u);
    const bool use_noise = (flags & SFK_EP_NOISE) != 0;
    auto noise_at = [&](int hh, int ww) -> float {
      return (use_noise && !v_d2s && tw < a.TW && hh < a.out_h && ww < a.out_w) ? __ldg(a.noise + static_cast<long>(hh) * a.out_w + ww) : 0.f;   // raw: scaled at use, so nothing waits on this load here
    };
    float nz_next = tile < tiles_per_group ? noise_at(t_h * a.TH + th, t_w * a.TW + tw) : 0.f;
    float nz_next2 = (v_m2 && tile < tiles_per_group) ? noise_at(t_h * a.TH + th + 8, t_w * a.TW + tw) : 0.f;   // second M tile
    // depth-to-space output: one raw noise value per output phase of this thread's coarse pixel
    const bool d2s_noise = use_noise && v_d2s;
    float nq0 = 0.f, nq1 = 0.f, nq2 = 0.f, nq3 = 0.f, nr0 = 0.f, nr1 = 0.f, nr2 = 0.f, nr3 = 0.f;   // nr*: second M tile
    auto noise4_at = [&](int hh, int ww) {
      if (d2s_noise && tw < a.TW && hh < a.out_h && ww < a.out_w) {
        const float2* r0 = reinterpret_cast<const float2*>(a.noise + (2L * hh) * (2 * a.out_w) + 2 * ww);
        const float2 u = __ldg(r0), l = __ldg(r0 + a.out_w);
        nq0 = u.x; nq1 = u.y; nq2 = l.x; nq3 = l.y;
      }
      if (d2s_noise && v_m2 && tw < a.TW && hh + 8 < a.out_h && ww < a.out_w) {
        const float2* r0 = reinterpret_cast<const float2*>(a.noise + (2L * (hh + 8)) * (2 * a.out_w) + 2 * ww);
        const float2 u = __ldg(r0), l = __ldg(r0 + a.out_w);
        nr0 = u.x; nr1 = u.y; nr2 = l.x; nr3 = l.y;
      }
    };
    if (tile < tiles_per_group) noise4_at(t_h * a.TH + th, t_w * a.TW + tw);
    const bool acc2 = a.acc_stages == 2;   // (acc_stages is 1 or 2: stage / phase without a division by a run-time value)
    for (int it = 0; tile < tiles_per_group; ++it) {
      const int as = acc2 ? (it & 1) : 0;
      const uint32_t aph = static_cast<uint32_t>(acc2 ? (it >> 1) : it) & 1u;
      int h = t_h * a.TH + th;
      const int w = t_w * a.TW + tw;
      bool valid = (tw < a.TW) && (h < a.out_h) && (w < a.out_w);
      const float nz_raw = nz_next, nz_raw2 = nz_next2;
      float nz4[4] = {nq0, nq1, nq2, nq3};
      const float nz4b[4] = {nr0, nr1, nr2, nr3};
      tile += gridDim.x;
      t_w += gridDim.x;
      while (t_w >= a.tiles_w) {
        t_w -= a.tiles_w;
        ++t_h;
      }
      if (tile < tiles_per_group) {
        nz_next = noise_at(t_h * a.TH + th, t_w * a.TW + tw);
        if (v_m2) nz_next2 = noise_at(t_h * a.TH + th + 8, t_w * a.TW + tw);
        noise4_at(t_h * a.TH + th, t_w * a.TW + tw);
      }
      const long long te0 = prof ? clock64() : 0;
      const bool ok = mbar_wait(&tmem_full_bar[as], aph, a.err);
      if (prof) t_we += clock64() - te0;
      valid = valid && ok;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float nz = a.noise_w * nz_raw;
      const int h_w0 = h - lrow, w_0 = w - tw;   // first pixel of this warp's box (staged store)
      const uint32_t tile_cols = static_cast<uint32_t>(m2n * a.num_acc * a.block_n);
      const int ksplit = kF32 ? a.ksplit : 1;
      uint32_t acc_col = static_cast<uint32_t>(as * ksplit) * tile_cols;   // first TMEM column of the tile (half)
      // NC = 16 or 32 accumulator columns per step (32 whenever block_n allows: twice the independent work per TMEM round trip)
      auto do_cols = [&](auto nc_tag, int acc, int c0, long pix, bool last_chunk) {
        constexpr int NC = decltype(nc_tag)::value;
        float v[NC], x[NC];
        long off = pix * a.out_c + n0 + c0;
        if (v_d2s) {   // this column block is one output phase: pixel (2h + ph/2, 2w + ph%2) of the fine grid
          const int ph = (n0 + c0) >> a.cq_log2, ch = (n0 + c0) & ((1 << a.cq_log2) - 1);   // Cq is a power of two (validate)
          off = (((static_cast<long>(n) * 2 * a.out_h + 2 * h + (ph >> 1)) * (2 * a.out_w) + 2 * w + (ph & 1)) << a.cq_log2) + ch;
          nz = a.noise_w * (ph == 0 ? nz4[0] : ph == 1 ? nz4[1] : ph == 2 ? nz4[2] : nz4[3]);
        }
        if (flags & (SFK_EP_XMASK | SFK_EP_GSDOT)) {   // start the activation load before the TMEM read completes
          if (valid) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) load8(xinp + off + i, x + i);
          } else {
#pragma unroll
            for (int i = 0; i < NC; ++i) x[i] = 0.f;
          }
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               acc_col + static_cast<uint32_t>((v_mg ? a.acc_blk[acc] : acc) * a.block_n + c0);
        if (NC == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        if constexpr (kF32) {   // sum the partial accumulators (fp32 adds, round to nearest)
          for (int p = 1; p < ksplit; ++p) {
            float u[NC];
            if (NC == 32) tmem_ld32(taddr + static_cast<uint32_t>(p) * tile_cols, u); else tmem_ld16(taddr + static_cast<uint32_t>(p) * tile_cols, u);
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] += u[i];
          }
        }
        if (last_chunk && a.early) {   // the accumulator is in registers: hand the TMEM stage back before the arithmetic and the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
        if (!(flags & SFK_EP_DSCALE) && (flags & (SFK_EP_NOISE | SFK_EP_BIAS))) {
          const float4* cb_ = reinterpret_cast<const float4*>(col_bias + c0);
          const float2 nz2 = make_float2(nz, nz);
#pragma unroll
          for (int i = 0; i < NC / 4; ++i) {   // packed fp32 (add.f32x2): half the issue slots of this issue/latency-bound loop
            const float4 bb = (flags & SFK_EP_BIAS) ? cb_[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float2 r0 = __fadd2_rn(make_float2(v[4 * i + 0], v[4 * i + 1]), __fadd2_rn(nz2, make_float2(bb.x, bb.y)));
            const float2 r1 = __fadd2_rn(make_float2(v[4 * i + 2], v[4 * i + 3]), __fadd2_rn(nz2, make_float2(bb.z, bb.w)));
            v[4 * i + 0] = r0.x;
            v[4 * i + 1] = r0.y;
            v[4 * i + 2] = r1.x;
            v[4 * i + 3] = r1.y;
          }
        } else if (flags & (SFK_EP_DSCALE | SFK_EP_NOISE | SFK_EP_BIAS)) {
          const float4* cd = reinterpret_cast<const float4*>(col_dscale + c0);
          const float4* cb_ = reinterpret_cast<const float4*>(col_bias + c0);
#pragma unroll
          for (int i = 0; i < NC / 4; ++i) {
            const float4 dd = cd[i], bb = cb_[i];
            v[4 * i + 0] = fmaf(v[4 * i + 0], dd.x, nz + bb.x);
            v[4 * i + 1] = fmaf(v[4 * i + 1], dd.y, nz + bb.y);
            v[4 * i + 2] = fmaf(v[4 * i + 2], dd.z, nz + bb.z);
            v[4 * i + 3] = fmaf(v[4 * i + 3], dd.w, nz + bb.w);
          }
        }
        if (flags & SFK_EP_RELU) {
#pragma unroll
          for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (flags & SFK_EP_LRELU) {
#pragma unroll
          for (int i = 0; i < NC; ++i) v[i] = lrelu_fwd(v[i]);
        }
        if (flags & SFK_EP_LRELU_RAW) {
          const float2 k02 = make_float2(0.2f, 0.2f);
#pragma unroll
          for (int i = 0; i < NC; i += 2) {
            const float2 m = __fmul2_rn(make_float2(v[i], v[i + 1]), k02);
            v[i] = fmaxf(v[i], m.x);
            v[i + 1] = fmaxf(v[i + 1], m.y);
          }
        }
        if ((flags & SFK_EP_GSDOT) && reg_gs) {
          // c0 is 0 or 32 here (block_n <= 64, NC == 32 or the 16-wide path with c0 in {0,16,32,48})
#pragma unroll
          for (int base = 0; base < kRegGs; base += NC) {
            if (base == c0) {
#pragma unroll
              for (int i = 0; i < NC; ++i) gsr[base + i] = fmaf(x[i], v[i], gsr[base + i]);
            }
          }
        } else if (flags & SFK_EP_GSDOT) {
#pragma unroll
          for (int hlf = 0; hlf < NC / 16; ++hlf) {
            float gsd[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) gsd[i] = x[hlf * 16 + i] * v[hlf * 16 + i];
            const float tot = warp_colsum16(gsd, lane);
            if (lane < 16) atomicAdd(&gs_acc[c0 + hlf * 16 + mycol], tot);
          }
        }
        if (flags & SFK_EP_XMASK) {
#pragma unroll
          for (int i = 0; i < NC; ++i) v[i] = x[i] > 0.f ? v[i] : 0.f;
        }
        if (flags & SFK_EP_COLSCALE) {
          const float4* cs = reinterpret_cast<const float4*>(col_scale + c0);
#pragma unroll
          for (int i = 0; i < NC / 4; ++i) {
            const float4 ss = cs[i];
            v[4 * i + 0] *= ss.x;
            v[4 * i + 1] *= ss.y;
            v[4 * i + 2] *= ss.z;
            v[4 * i + 3] *= ss.w;
          }
        }
        if (!kF32 && v_ts) {
          // stage the bf16 rows in shared memory (swizzled like the tensor map) and let TMA write full lines: a direct
          // 16-byte store per thread touches 16-32 different 128-byte lines per warp instruction, and those LSU wavefronts
          // share the L1/shared-memory data pipe with the tensor core's operand reads
          const int co = c0 % a.ts_slabw;
          if (co == 0) {   // about to refill a buffer: its previous store must have finished reading it
            if (lane == 0) {
              if (a.ts_nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
          }
          const uint32_t buf = ts_base + static_cast<uint32_t>(ts_buf * 32 * a.ts_slabw * 2);
          if (tw < a.TW) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
              const uint32_t j = static_cast<uint32_t>((co + i) >> 3);
              sts128(buf + ts_row + ((j ^ ts_swz) << 4), pack8(v + i));
            }
          }
          if (co + NC == a.ts_slabw) {   // slab complete
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              const int cs = n0 + c0 + NC - a.ts_slabw;
              if (v_d2s)
                tma_store_5d(&a.mapO, buf, cs & ((2 << a.cq_log2) - 1), w_0, cs >> (a.cq_log2 + 1), h_w0, n);
              else
                tma_store_5d(&a.mapO, buf, cs, w_0, 0, h_w0, n * a.num_acc + acc);
            }
            ts_buf = (ts_buf + 1) % a.ts_nbuf;
          }
        } else if (NC == 32 && !kF32 && v_xs && !(flags & SFK_EP_ACCUM)) {
          // 4x4 transpose of the 16-byte pieces across the four lanes of a pixel quad (two butterfly exchanges): afterwards lane
          // i of a quad owns piece i of all four pixels, so one store instruction writes 64 contiguous bytes per quad and touches
          // 8 lines per warp instead of 16 (32 for the depth-to-space output)
          const uint4 I0 = pack8(v), I1 = pack8(v + 8), I2 = pack8(v + 16), I3 = pack8(v + 24);
          const int qi = lane & 3;
          const bool hi = (qi & 2) != 0, odd = (qi & 1) != 0;
          const uint4 r0 = shfl_xor_u4(hi ? I0 : I2, 2), r1 = shfl_xor_u4(hi ? I1 : I3, 2);
          const uint4 J0 = hi ? r0 : I0, J1 = hi ? r1 : I1, J2 = hi ? I2 : r0, J3 = hi ? I3 : r1;
          const uint4 ra = shfl_xor_u4(odd ? J0 : J1, 1), rb = shfl_xor_u4(odd ? J2 : J3, 1);
          const uint4 K[4] = {odd ? ra : J0, odd ? J1 : ra, odd ? rb : J2, odd ? J3 : rb};
          const unsigned vmask = __ballot_sync(0xffffffffu, valid);
          const long pitch = v_d2s ? (a.out_c >> 1) : a.out_c;   // elements between the pixels of neighbouring lanes
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if ((vmask >> (lane - qi + k)) & 1u) stg8(outp + off + (k - qi) * pitch + qi * 8, K[k]);
          }
        } else if (valid) {
          if (flags & SFK_EP_ACCUM) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
              float o[8];
              load8p(outp + off + i, o);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[i + e] += o[e];
            }
          }
#pragma unroll
          for (int i = 0; i < NC; i += 8) store8(outp + off + i, v + i);
        }
      };
      for (int half = 0; half < m2n; ++half) {
        if (half == 1) {   // second M tile of the stage: output rows 8 further down, the next accumulator block
          h += 8;
          valid = ok && (tw < a.TW) && (h < a.out_h) && (w < a.out_w);
          nz = a.noise_w * nz_raw2;
#pragma unroll
          for (int p4 = 0; p4 < 4; ++p4) nz4[p4] = nz4b[p4];
          acc_col += static_cast<uint32_t>(a.num_acc * a.block_n);
        }
        for (int acc = 0; acc < a.num_acc; ++acc) {
          const long pix = ((static_cast<long>(n) * a.num_acc + acc) * a.out_h + h) * a.out_w + w;
          const bool last_acc = half == m2n - 1 && acc == a.num_acc - 1;
          if ((a.block_n & 31) == 0 && (!v_d2s || ((a.out_c >> 2) & 31) == 0)) {
            for (int c0 = 0; c0 < a.block_n; c0 += 32) do_cols(std::integral_constant<int, 32>{}, acc, c0, pix, last_acc && c0 + 32 >= a.block_n);
          } else {
            for (int c0 = 0; c0 < a.block_n; c0 += 16) do_cols(std::integral_constant<int, 16>{}, acc, c0, pix, last_acc && c0 + 16 >= a.block_n);
          }
        }
      }
      if (!a.early) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);  // accumulator stage drained (one arrival per warp)
      }
    }
    if (v_ts && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
    if (reg_gs) {
#pragma unroll
      for (int i = 0; i < kRegGs; ++i) {
        if (i < a.block_n) {
          const float tot = warp_sum(gsr[i]);
          if (lane == 0) atomicAdd(&gs_acc[i], tot);
        }
      }
    }
    if (prof) {
      atomicAdd(&g_role_cycles[5], static_cast<unsigned long long>(t_we));
      atomicAdd(&g_role_cycles[6], static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(&g_role_cycles[7], static_cast<unsigned long long>((tiles_per_group - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)));
    }
  }
  __syncthreads();
  if ((a.flags & SFK_EP_GSDOT) && threadIdx.x < a.block_n) {
    atomicAdd(a.gs + static_cast<long>(n) * a.vec_stride + n0 + threadIdx.x, gs_acc[threadIdx.x]);
  }
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// CUDA-core cross-check with the identical contract (one thread = one position x 16 channels).
template <typename T>
struct RefArgs {
  const T* A;
  const T* B;
  int n_img, a_h, a_w, a_c, a_planes, b_rows, b_per_sample;
  int out_h, out_w, out_c, num_acc, block_n, num_taps, flags;
  int vec_stride;
  int out_d2s, a_s2d;
  float noise_w;
  T* out;
  const float* dscale;
  const float* bias;
  const float* noise;
  const T* xin;
  const float* colscale;
  float* gs;
  KTap taps[SFK_MAX_TAPS];
};

template <typename T>
__global__ void igemm_ref_kernel(const __grid_constant__ RefArgs<T> a) {
  const int chunks_per_pos = a.out_c / 16;
  const long total = static_cast<long>(a.n_img) * a.num_acc * a.out_h * a.out_w * chunks_per_pos;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ch = idx % chunks_per_pos;
    long r = idx / chunks_per_pos;
    const int w = r % a.out_w;
    r /= a.out_w;
    const int h = r % a.out_h;
    r /= a.out_h;
    const int acc = r % a.num_acc;
    const int n = r / a.num_acc;
    const int col0 = ch * 16;
    const int nblk = col0 / a.block_n;            // which weight-row block this channel group lives in
    const int cin_blk = col0 - nblk * a.block_n;  // offset inside it
    float v[16], gsd[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    for (int t = 0; t < a.num_taps; ++t) {
      const KTap& tp = a.taps[t];
      if (tp.acc != acc) continue;
      const int ih = h + tp.dy, iw = w + tp.dx;
      if (ih < 0 || ih >= a.a_h || iw < 0 || iw >= a.a_w) continue;
      const T* ap = a.A + ((((static_cast<long>(n) * a.a_planes + tp.plane) * a.a_h + ih) * a.a_w + iw) * a.a_c);
      const int Cq = a.a_c / 4;
      const long brow0 = static_cast<long>(a.b_per_sample ? n : 0) * a.b_rows + tp.brow + nblk * a.block_n + cin_blk;
      for (int k = 0; k < a.a_c; ++k) {
        float av;
        if (a.a_s2d) {   // K index = phase*Cq + c of fine pixel (2ih + phase/2, 2iw + phase%2)
          const int ph = k / Cq, c = k % Cq;
          av = to_f32(a.A[((static_cast<long>(n) * 2 * a.a_h + 2 * ih + (ph >> 1)) * (2 * a.a_w) + 2 * iw + (ph & 1)) * Cq + c]);
        } else {
          av = to_f32(ap[k]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const long br = brow0 + i;
          const float bv = (tp.brow + nblk * a.block_n + cin_blk + i < a.b_rows) ? to_f32(a.B[br * a.a_c + k]) : 0.f;
          v[i] += av * bv;
        }
      }
    }
    epilogue16(a, n, acc, h, w, col0, true, v, gsd);
    if (a.flags & SFK_EP_GSDOT) {
#pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(a.gs + static_cast<long>(n) * a.vec_stride + col0 + i, gsd[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int validate(const sfk_igemm_desc* d) {
  SFK_REQUIRE(d != nullptr, SFK_E_ARG, "igemm: null descriptor");
  SFK_REQUIRE(d->a && d->b && d->out, SFK_E_ARG, "igemm: null operand");
  SFK_REQUIRE(sfk_aligned16(d->a) && sfk_aligned16(d->b) && sfk_aligned16(d->out), SFK_E_ALIGN, "igemm: operands must be 16B aligned");
  SFK_REQUIRE(d->a_c >= 16 && d->a_c % 16 == 0, SFK_E_SHAPE, "igemm: Cin must be a multiple of 16");
  SFK_REQUIRE(d->block_n >= 16 && d->block_n % 16 == 0 && d->block_n <= 256, SFK_E_SHAPE, "igemm: block_n must be 16..256, multiple of 16");
  SFK_REQUIRE(d->num_acc >= 1 && d->num_acc * d->block_n <= 512, SFK_E_SHAPE, "igemm: num_acc*block_n must fit 512 TMEM columns");
  SFK_REQUIRE(d->out_c % d->block_n == 0, SFK_E_SHAPE, "igemm: out_c must be a multiple of block_n");
  SFK_REQUIRE(d->num_taps >= 1 && d->num_taps <= SFK_MAX_TAPS, SFK_E_SHAPE, "igemm: bad tap count");
  SFK_REQUIRE(d->n_img >= 1 && d->out_h >= 1 && d->out_w >= 1 && d->a_h >= 1 && d->a_w >= 1 && d->a_planes >= 1, SFK_E_SHAPE, "igemm: bad dims");
  SFK_REQUIRE(d->b_samples == 1 || d->b_samples == d->n_img, SFK_E_SHAPE, "igemm: b_samples must be 1 or n_img");
  for (int t = 0; t < d->num_taps; ++t) {
    SFK_REQUIRE(d->taps[t].acc >= 0 && d->taps[t].acc < d->num_acc, SFK_E_SHAPE, "igemm: tap accumulator out of range");
    SFK_REQUIRE(d->taps[t].plane >= 0 && d->taps[t].plane < d->a_planes, SFK_E_SHAPE, "igemm: tap plane out of range");
    SFK_REQUIRE(d->taps[t].brow >= 0 && d->taps[t].brow + d->out_c <= d->b_rows, SFK_E_SHAPE, "igemm: tap weight rows out of range");
  }
  if (d->out_d2s) SFK_REQUIRE(d->num_acc == 1 && d->out_c % 64 == 0 && (d->out_c & (d->out_c - 1)) == 0 && !(d->flags & (SFK_EP_XMASK | SFK_EP_GSDOT | SFK_EP_ACCUM | SFK_EP_COLSCALE)),
                              SFK_E_SHAPE, "igemm: depth-to-space output needs one accumulator, out_c a power of two >= 64 and a forward epilogue");
  if (d->a_s2d) SFK_REQUIRE(d->a_planes == 1 && d->a_c % 64 == 0, SFK_E_SHAPE, "igemm: space-to-depth input needs a_c % 64 == 0");
  if (d->flags & SFK_EP_DSCALE) SFK_REQUIRE(d->dscale, SFK_E_ARG, "igemm: dscale missing");
  if (d->flags & SFK_EP_BIAS) SFK_REQUIRE(d->bias, SFK_E_ARG, "igemm: bias missing");
  if (d->flags & SFK_EP_NOISE) SFK_REQUIRE(d->noise, SFK_E_ARG, "igemm: noise missing");
  if (d->flags & (SFK_EP_XMASK | SFK_EP_GSDOT)) SFK_REQUIRE(d->xin && sfk_aligned16(d->xin), SFK_E_ARG, "igemm: xin missing");
  if (d->flags & SFK_EP_GSDOT) SFK_REQUIRE(d->gs, SFK_E_ARG, "igemm: gs missing");
  if (d->flags & SFK_EP_COLSCALE) SFK_REQUIRE(d->colscale, SFK_E_ARG, "igemm: colscale missing");
  return 0;
}

void fill_taps(const sfk_igemm_desc* d, KTap* taps) {
  bool seen[32] = {false};
  for (int t = 0; t < d->num_taps; ++t) {
    taps[t].dy = d->taps[t].dy;
    taps[t].dx = d->taps[t].dx;
    taps[t].plane = d->taps[t].plane;
    taps[t].acc = d->taps[t].acc;
    taps[t].brow = d->taps[t].brow;
    taps[t].first = seen[d->taps[t].acc] ? 0 : 1;
    seen[d->taps[t].acc] = true;
  }
}

}  // namespace

extern "C" int sfk_igemm_v1(const sfk_igemm_desc* d, sfk_stream_t stream) {
  int rc = validate(d);
  if (rc) return rc;
  SFK_REQUIRE(!sfk_act_f32(), SFK_E_ARG, "igemm_v1: bf16 activations only");
  SFK_REQUIRE(!d->out_d2s && !d->a_s2d, SFK_E_ARG, "igemm_v1: fused resampling is only implemented by sfk_igemm / sfk_igemm_ref");
  EncodeTiledFn enc = get_encode_fn();
  SFK_REQUIRE(enc != nullptr, SFK_E_DRIVER, "igemm: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");

  IgemmKArgs k;
  memset(&k, 0, sizeof(k));
  const int KC = (d->a_c % 64 == 0) ? 64 : (d->a_c % 32 == 0 ? 32 : 16);
  k.KC = KC;
  k.num_cblk = d->a_c / KC;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  k.layout_type = KC == 64 ? 2 : (KC == 32 ? 4 : 6);
  k.sbo_bytes = 8 * KC * 2;
  k.TW = d->out_w > 8 ? 16 : (d->out_w > 4 ? 8 : 4);
  k.TH = 128 / k.TW;
  k.tiles_w = (d->out_w + k.TW - 1) / k.TW;
  k.tiles_h = (d->out_h + k.TH - 1) / k.TH;
  k.n_img = d->n_img;
  k.out_h = d->out_h;
  k.out_w = d->out_w;
  k.out_c = d->out_c;
  k.block_n = d->block_n;
  k.num_acc = d->num_acc;
  k.num_taps = d->num_taps;
  k.b_per_sample = d->b_samples > 1 ? 1 : 0;
  k.a_stage_bytes = 128 * KC * 2;
  k.b_stage_bytes = ((d->block_n * KC * 2 + 1023) / 1024) * 1024;
  const int cols = d->num_acc * d->block_n;
  k.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  k.flags = d->flags;
  k.vec_stride = d->vec_stride > 0 ? d->vec_stride : d->out_c;
  k.noise_w = d->noise_w;
  k.out = static_cast<__nv_bfloat16*>(d->out);
  k.dscale = d->dscale;
  k.bias = d->bias;
  k.noise = d->noise;
  k.xin = static_cast<const __nv_bfloat16*>(d->xin);
  k.colscale = d->colscale;
  k.gs = d->gs;
  k.err = d->err;
  fill_taps(d, k.taps);
  const int num_k = k.num_cblk * k.num_taps;
  int stages = d->stages;
  if (stages <= 0) {
    const int budget = 100 * 1024;  // two CTAs per SM
    stages = budget / (k.a_stage_bytes + k.b_stage_bytes);
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > num_k) stages = num_k;
  if (stages < 1) stages = 1;
  k.stages = stages;

  // A: [n][planes][h][w][c] bf16, box {KC, TW, TH, 1, 1}
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->a_c, (cuuint64_t)d->a_w, (cuuint64_t)d->a_h, (cuuint64_t)d->a_planes, (cuuint64_t)d->n_img};
    cuuint64_t strides[4] = {(cuuint64_t)d->a_c * 2, (cuuint64_t)d->a_w * d->a_c * 2, (cuuint64_t)d->a_h * d->a_w * d->a_c * 2,
                             (cuuint64_t)d->a_planes * d->a_h * d->a_w * d->a_c * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)k.TW, (cuuint32_t)k.TH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&k.mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->a), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SFK_REQUIRE(r == CUDA_SUCCESS, SFK_E_DRIVER, "igemm: cuTensorMapEncodeTiled(A) failed");
  }
  // B: [samples][rows][c] bf16, box {KC, block_n, 1}
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->a_c, (cuuint64_t)d->b_rows, (cuuint64_t)d->b_samples};
    cuuint64_t strides[2] = {(cuuint64_t)d->a_c * 2, (cuuint64_t)d->b_rows * d->a_c * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)d->block_n, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&k.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->b), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SFK_REQUIRE(r == CUDA_SUCCESS, SFK_E_DRIVER, "igemm: cuTensorMapEncodeTiled(B) failed");
  }
  const size_t smem = static_cast<size_t>(stages) * (k.a_stage_bytes + k.b_stage_bytes) + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    smem_set = 200 * 1024;
  }
  dim3 grid(static_cast<unsigned>(k.tiles_w * k.tiles_h * d->n_img), static_cast<unsigned>(d->out_c / d->block_n));
  igemm_tc_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(k);
  return sfk_check_launch("igemm_tc_kernel");
}

namespace {
// Every size below is in BYTES of the storage element (es = 2 bf16 / 4 fp32), so one code path plans both the kind::f16 and the
// kind::tf32 launches.  `base`/`n_mult`: the split-tf32 mode reads hi/lo copies from the workspace, lo n_img images further on.
int encode_a_map(EncodeTiledFn enc, CUtensorMap* map, const sfk_igemm_desc* d, const void* base, int n_mult, int es, int KC, int TW,
                 int rows, CUtensorMapSwizzle swz) {
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const cuuint64_t e = (cuuint64_t)es;
  if (d->a_s2d) {   // [n][2*a_h][2*a_w][Cq] viewed as {pixel pair (2Cq), W, row phase, H, N}
    const cuuint64_t cq = (cuuint64_t)d->a_c / 4, frow = 2 * (cuuint64_t)d->a_w * cq * e;   // bytes of one fine row
    cuuint64_t dims[5] = {2 * cq, (cuuint64_t)d->a_w, 2, (cuuint64_t)d->a_h, (cuuint64_t)d->n_img * n_mult};
    cuuint64_t strides[4] = {2 * cq * e, frow, 2 * frow, 2 * (cuuint64_t)d->a_h * frow};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)TW, 1, (cuuint32_t)rows, 1};
    cuuint32_t es1[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, dt, 5, const_cast<void*>(base), dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
  }
  cuuint64_t dims[5] = {(cuuint64_t)d->a_c, (cuuint64_t)d->a_w, (cuuint64_t)d->a_h, (cuuint64_t)d->a_planes, (cuuint64_t)d->n_img * n_mult};
  cuuint64_t strides[4] = {(cuuint64_t)d->a_c * e, (cuuint64_t)d->a_w * d->a_c * e, (cuuint64_t)d->a_h * d->a_w * d->a_c * e,
                           (cuuint64_t)d->a_planes * d->a_h * d->a_w * d->a_c * e};
  cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)TW, (cuuint32_t)rows, 1, 1};
  cuuint32_t es1[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, dt, 5, const_cast<void*>(base), dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// hi = x rounded to tf32 (10-bit mantissa, round to nearest), lo = tf32(x - hi): x*y ~= hi_x*hi_y + lo_x*hi_y + hi_x*lo_y with a
// relative error of ~2^-21 per product instead of tf32's 2^-11 (the tensor core truncates fp32 bit patterns to tf32; hi and lo are
// exactly representable, so that truncation is the identity).
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t u = __float_as_uint(v);
  u = (u + 0x1000u) & 0xffffe000u;
  return __uint_as_float(u);
}
__global__ void tf32_split_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, long n4) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(x + i);
    float4 h, l;
    h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
    l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
    hi[i] = h;
    lo[i] = l;
  }
}

struct IgemmPlan {
  Igemm2Args k;
  sfk_igemm_desc desc;     // copy (CUDA-core mode and diagnostics)
  dim3 grid;
  size_t smem;
  int fct;                 // epilogue flag set the launch is specialised for
  int var;                 // launch variant bits (kVar*), -1 if the staged TMA store is on (run-time variant only)
  int f32;                 // storage: 0 bf16, 1 fp32
  int use_ref;             // fp32 storage with conv math 3: CUDA-core kernel
  // split-tf32 operand copies (workspace): [A.hi][A.lo][B.hi][B.lo]
  const float* a_src; const float* b_src;
  float* a_hi; float* b_hi;
  long a_n4, b_n4;
};

size_t split_ws_bytes(const sfk_igemm_desc* d) {
  const size_t na = static_cast<size_t>(d->n_img) * d->a_planes * d->a_h * d->a_w * d->a_c * (d->a_s2d ? 1 : 1);
  const size_t nb = static_cast<size_t>(d->b_samples) * d->b_rows * d->a_c;
  return 2 * (na + nb) * sizeof(float);
}

int g_conv_math = 0;   // fp32 storage: 0/2 = split tf32 (three passes), 1 = plain tf32, 3 = CUDA cores

// Instantiated (epilogue, variant) pairs: every launch of the bf16 product path hits a fully specialised kernel; anything else
// (and the whole fp32 parity mode) runs the same source with run-time flags.
template <typename T>
const void* kernel_for(int fct, int var) {
  if (!std::is_same<T, float>::value && var >= 0) {
    switch ((fct << 4) | var) {
#define SFK_CASE(F, V) case (((F) << 4) | (V)): return reinterpret_cast<const void*>(&igemm_tc2_kernel<F, __nv_bfloat16, V>)
      SFK_CASE(0, 0); SFK_CASE(0, kVarM2); SFK_CASE(0, kVarXS); SFK_CASE(0, kVarXS | kVarM2); SFK_CASE(0, kVarMG); SFK_CASE(0, kVarXS | kVarMG); SFK_CASE(0, kVarM2 | kVarMG); SFK_CASE(0, kVarXS | kVarM2 | kVarMG);
      SFK_CASE(SFK_EP_BIAS | SFK_EP_RELU, 0); SFK_CASE(SFK_EP_BIAS | SFK_EP_RELU, kVarM2);
      SFK_CASE(SFK_EP_NOISE | SFK_EP_BIAS | SFK_EP_LRELU_RAW, 0); SFK_CASE(SFK_EP_NOISE | SFK_EP_BIAS | SFK_EP_LRELU_RAW, kVarM2);
      SFK_CASE(SFK_EP_NOISE | SFK_EP_BIAS | SFK_EP_LRELU_RAW, kVarD2S); SFK_CASE(SFK_EP_NOISE | SFK_EP_BIAS | SFK_EP_LRELU_RAW, kVarD2S | kVarXS);
      SFK_CASE(SFK_EP_XMASK, 0); SFK_CASE(SFK_EP_XMASK, kVarM2);
      SFK_CASE(SFK_EP_GSDOT | SFK_EP_COLSCALE, 0); SFK_CASE(SFK_EP_GSDOT | SFK_EP_COLSCALE, kVarM2);
#undef SFK_CASE
      default: break;
    }
  }
  switch (fct) {
#define SFK_CASE(F) case (F): return reinterpret_cast<const void*>(&igemm_tc2_kernel<F, T, -1>)
    SFK_CASE(0);
    SFK_CASE(SFK_EP_BIAS | SFK_EP_RELU);
    SFK_CASE(SFK_EP_NOISE | SFK_EP_BIAS | SFK_EP_LRELU_RAW);
    SFK_CASE(SFK_EP_XMASK);
    SFK_CASE(SFK_EP_GSDOT | SFK_EP_COLSCALE);
#undef SFK_CASE
    default: return reinterpret_cast<const void*>(&igemm_tc2_kernel<-1, T, -1>);
  }
}

int build_plan(const sfk_igemm_desc* d, IgemmPlan* P) {
  int rc = validate(d);
  if (rc) return rc;
  memset(P, 0, sizeof(*P));
  P->desc = *d;
  P->f32 = sfk_act_f32();
  if (P->f32 && g_conv_math == 3) {
    P->use_ref = 1;
    return 0;
  }
  EncodeTiledFn enc = get_encode_fn();
  SFK_REQUIRE(enc != nullptr, SFK_E_DRIVER, "igemm: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  Igemm2Args& k = P->k;
  KGroup hg[kMaxGroups];   // host-side view of the load groups; the kernel's part is copied into k.groups at the end
  memset(hg, 0, sizeof(hg));
  const int es = P->f32 ? 4 : 2;
  k.passes = (P->f32 && g_conv_math != 1) ? 3 : 1;
  k.b_samples = d->b_samples;
  const void* a_base = d->a;
  const void* b_base = d->b;
  if (k.passes == 3) {
    SFK_REQUIRE(d->ws != nullptr && d->ws_bytes >= split_ws_bytes(d) && sfk_aligned16(d->ws), SFK_E_ARG,
                "igemm: split-tf32 mode needs desc.ws >= sfk_igemm_workspace_bytes()");
    const size_t na = static_cast<size_t>(d->n_img) * d->a_planes * d->a_h * d->a_w * d->a_c;
    const size_t nb = static_cast<size_t>(d->b_samples) * d->b_rows * d->a_c;
    P->a_src = static_cast<const float*>(d->a);
    P->b_src = static_cast<const float*>(d->b);
    P->a_hi = static_cast<float*>(d->ws);
    P->b_hi = P->a_hi + 2 * na;
    P->a_n4 = static_cast<long>(na / 4);
    P->b_n4 = static_cast<long>(nb / 4);
    a_base = P->a_hi;
    b_base = P->b_hi;
  }
  const int kdim = d->a_s2d ? d->a_c / 2 : d->a_c;   // contiguous K extent in memory (space-to-depth: one pixel pair = 2*Cq)
  const int row_bytes = (kdim * es) % 128 == 0 ? 128 : ((kdim * es) % 64 == 0 ? 64 : 32);
  const int KC = row_bytes / es;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  k.KC = KC;
  k.num_cblk = d->a_c / KC;
  k.out_d2s = d->out_d2s;
  k.a_s2d = d->a_s2d;
  k.cpa = d->a_s2d ? kdim / KC : 1;
  for (k.cq_log2 = 0; (4 << k.cq_log2) < d->out_c; ++k.cq_log2) {}
  k.row_bytes = row_bytes;
  k.layout_type = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  k.sbo_bytes = 8 * k.row_bytes;
  k.TW = d->out_w > 8 ? 16 : (d->out_w > 4 ? 8 : 4);
  k.TH = 128 / k.TW;
  k.tiles_w = (d->out_w + k.TW - 1) / k.TW;
  k.tiles_h = (d->out_h + k.TH - 1) / k.TH;
  k.n_img = d->n_img; k.out_h = d->out_h; k.out_w = d->out_w; k.out_c = d->out_c;
  k.block_n = d->block_n; k.num_acc = d->num_acc; k.num_taps = d->num_taps;
  k.n_blocks = d->out_c / d->block_n;
  k.b_per_sample = d->b_samples > 1 ? 1 : 0;
  k.flags = d->flags;
  k.vec_stride = d->vec_stride > 0 ? d->vec_stride : d->out_c;
  k.noise_w = d->noise_w;
  k.out = d->out;
  k.dscale = d->dscale; k.bias = d->bias; k.noise = d->noise;
  k.xin = d->xin;
  k.colscale = d->colscale; k.gs = d->gs; k.err = d->err;

  // ---- group taps that can share ONE activation load
  //   halo mode (row pitch 16): all taps of a plane whose shifts span <= 2 in x and y read one (TH+span_y[+1]) x 16 box;
  //                             a tap enters it dy*16+dx rows down (the swizzle is a function of the absolute smem address,
  //                             so an operand may start at any row; the descriptor's base-offset field stays 0);
  //                             the tile keeps TW = 16 - span_x useful columns
  //   dy mode (otherwise):      taps that differ only in dy share a box TH+span rows tall (row offsets are multiples of 8)
  // Halo mode pays off where the layer is bound by activation loads (small channel counts: the whole weight set is resident in
  // smem); for wide layers the 16->14 useful columns cost more tensor time than the saved loads.  SFK_HALO=0/1 forces it.
  static const int halo_env = env_int("SFK_HALO", -1);
  const int halves = k.passes == 3 ? 2 : 1;
  const int b_total_est = halves * (d->a_c / KC) * d->num_taps * (((d->block_n * row_bytes + 1023) / 1024) * 1024);
  // (measured: forward convs at 1024^2 / 512^2 gain 15-20 %; the data-gradient launches are bound by their heavier epilogue and
  //  lose ~8 % to the narrower tile, so they keep the dy-shared mode)
  const bool light_epilogue = (d->flags & (SFK_EP_GSDOT | SFK_EP_XMASK)) == 0;
  // the fused-resampling launches (4 phases of weights) keep their whole weight set resident at one CTA per SM
  static const int rl_env = env_int("SFK_S2D_RESIDENT", 0);   // measured at 1024^2: streamed weights at 2 CTAs/SM 455 us, resident at 1 CTA/SM 571 us
  const int resident_limit = (d->out_d2s || (d->a_s2d && rl_env)) ? 150 * 1024 : 72 * 1024;
  const bool can_reside = b_total_est <= resident_limit && halves * (d->a_c / KC) <= kMaxStages;   // one descriptor slot per resident channel block
  const bool halo = k.TW == 16 && !d->a_s2d && !d->out_d2s && (halo_env >= 0 ? (halo_env != 0 && can_reside) : (can_reside && light_epilogue));
  // Two M tiles per stage where the weights are streamed (they do not fit shared memory) in 128-column tiles: the stage's
  // weight tile (3 taps x 16 KB) then feeds 256 output rows instead of 128, which takes ~15 % off the shared-memory port
  // (TMA writes + tensor-core operand reads).  Needs both halves' accumulators double-buffered: 4 x 128 columns.
  static const int m2_env = env_int("SFK_M2", 1);
  // (fp32 storage with a long K: the TMEM columns go to partial accumulators instead, see ksplit below)
  // (The 4-accumulator transposed conv at 2 x 4 x 64 columns was tried too, SFK_M2_TCONV=1: its accumulators are then single-buffered
  //  and only the 17^2 layer gains (78 -> 67 us); 129^2 158 -> 171 us, 257^2 148 -> 181 us.  Off.)
  static const int m2t_env = env_int("SFK_M2_TCONV", 0);
  const bool m2_shape = (d->block_n == 128 && d->num_acc == 1) || (m2t_env && !P->f32 && d->num_acc == 4 && d->block_n == 64 && d->a_c >= 128);
  k.m2 = (m2_env && k.TW == 16 && !halo && !can_reside && m2_shape &&
          d->out_h >= 16 && row_bytes == 128 && !(P->f32 && d->a_c >= 512)) ? 1 : 0;
  if (k.m2) {
    k.TH = 16;
    k.tiles_h = (d->out_h + k.TH - 1) / k.TH;
  }
  const bool share = k.TW >= 8;
  int ng = 0;
  int dymin[kMaxGroups], dymax[kMaxGroups], dxmin[kMaxGroups], dxmax[kMaxGroups], tdy[kMaxGroups][kMaxGroupTaps], tdx[kMaxGroups][kMaxGroupTaps];
  for (int t = 0; t < d->num_taps; ++t) {
    const sfk_tap& tp = d->taps[t];
    int g = -1;
    if (share) {
      for (int q = 0; q < ng; ++q) {
        if (hg[q].plane != tp.plane || hg[q].ntaps >= kMaxGroupTaps) continue;
        if (!halo && dxmin[q] != tp.dx) continue;
        const int lo = tp.dy < dymin[q] ? tp.dy : dymin[q], hi = tp.dy > dymax[q] ? tp.dy : dymax[q];
        const int xlo = tp.dx < dxmin[q] ? tp.dx : dxmin[q], xhi = tp.dx > dxmax[q] ? tp.dx : dxmax[q];
        if (hi - lo <= 2 && xhi - xlo <= 2) { g = q; break; }
      }
    }
    if (g < 0) {
      SFK_REQUIRE(ng < kMaxGroups, SFK_E_SHAPE, "igemm: too many tap groups");
      g = ng++;
      hg[g].plane = tp.plane; hg[g].ntaps = 0;
      dymin[g] = dymax[g] = tp.dy; dxmin[g] = dxmax[g] = tp.dx;
    }
    KGroup& G = hg[g];
    if (tp.dy < dymin[g]) dymin[g] = tp.dy;
    if (tp.dy > dymax[g]) dymax[g] = tp.dy;
    if (tp.dx < dxmin[g]) dxmin[g] = tp.dx;
    if (tp.dx > dxmax[g]) dxmax[g] = tp.dx;
    tdy[g][G.ntaps] = tp.dy; tdx[g][G.ntaps] = tp.dx;
    G.acc[G.ntaps] = tp.acc; G.brow[G.ntaps] = tp.brow; G.bidx[G.ntaps] = t;
    G.ntaps++;
  }
  k.num_groups = ng;
  int max_xspan = 0;
  for (int g = 0; g < ng; ++g) if (dxmax[g] - dxmin[g] > max_xspan) max_xspan = dxmax[g] - dxmin[g];
  k.TWB = k.TW;
  k.TW = k.TWB - max_xspan;                      // useful columns per tile row
  k.tiles_w = (d->out_w + k.TW - 1) / k.TW;
  bool seen[32] = {false};
  int max_rows_extra = 0, max_gt = 0;
  for (int g = 0; g < ng; ++g) {
    KGroup& G = hg[g];
    G.dy_min = dymin[g];
    G.dx_min = dxmin[g];
    const int extra = (dymax[g] - dymin[g]) + ((dxmax[g] - dxmin[g]) > 0 ? 1 : 0);   // +1 row: the M index runs past the last row
    G.map = extra;
    G.bytes = (k.TH + extra) * k.TWB * k.row_bytes;
    if (extra > max_rows_extra) max_rows_extra = extra;
    if (G.ntaps > max_gt) max_gt = G.ntaps;
    for (int j = 0; j < G.ntaps; ++j) {
      G.roff[j] = (tdy[g][j] - dymin[g]) * k.TWB + (tdx[g][j] - dxmin[g]);
      G.first[j] = seen[G.acc[j]] ? 0 : 1;
      seen[G.acc[j]] = true;
    }
  }
  const int max_span = max_rows_extra;
  // ---- merged taps.  Taps of one load group with the same operand offset read the SAME activation rows; if their accumulators
  // occupy adjacent TMEM column blocks and their weight tiles are adjacent in shared memory they are ONE tcgen05.mma with a
  // multiple of block_n columns: the activation operand is then read from shared memory once instead of once per tap.  The
  // 4-phase transposed conv has 4 distinct shifts for its 9 taps (accumulator sets {0,1,2,3}, {0,1}, {0,2}, {0}): with the column
  // order (1,0,2,3) it becomes 4 MMAs of 256/128/128/64 columns -- 34 KB instead of 54 KB of operand reads per 16-deep k-slice,
  // against 288 tensor cycles (these launches were bound by the shared-memory port).
  int wm[kMaxGroups][kMaxGroupTaps];
  for (int g = 0; g < kMaxGroups; ++g)
    for (int j = 0; j < kMaxGroupTaps; ++j) wm[g][j] = 1;
  for (int i = 0; i < 4; ++i) k.acc_blk[i] = i;
  static const int merge_env = env_int("SFK_TAP_MERGE", 1);
  const int max_run = 256 / d->block_n;
  if (merge_env && d->num_acc > 1 && d->num_acc <= 4 && (d->block_n * k.row_bytes) % 1024 == 0 && max_run >= 2) {
    const int na = d->num_acc;
    int perm[4] = {0, 1, 2, 3}, best_perm[4] = {0, 1, 2, 3}, best = 1 << 30;
    // MMAs needed under a given accumulator -> column-block map (perm[acc] = block)
    auto count_mmas = [&](const int* pm) {
      int total = 0;
      for (int g = 0; g < ng; ++g) {
        const KGroup& G = hg[g];
        bool done[kMaxGroupTaps] = {false};
        for (int j = 0; j < G.ntaps; ++j) {
          if (done[j]) continue;
          bool used[4] = {false, false, false, false};   // column blocks of the taps sharing roff[j]
          int dup = 0;
          for (int q = j; q < G.ntaps; ++q)
            if (!done[q] && G.roff[q] == G.roff[j]) {
              done[q] = true;
              if (used[pm[G.acc[q]]]) ++dup; else used[pm[G.acc[q]]] = true;
            }
          int run = 0;
          for (int c = 0; c < na; ++c) {
            if (used[c]) { if (run == 0 || run == max_run) { ++total; run = 0; } ++run; } else run = 0;
          }
          total += dup;
        }
      }
      return total;
    };
    // all permutations of na <= 4 blocks (Heap's algorithm is overkill: enumerate base-na digits and keep the bijections)
    int lim = 1;
    for (int i = 0; i < na; ++i) lim *= na;
    for (int code = 0; code < lim; ++code) {
      int c = code, msk = 0;
      for (int i = 0; i < na; ++i) { perm[i] = c % na; c /= na; msk |= 1 << perm[i]; }
      if (msk != (1 << na) - 1) continue;
      const int cnt = count_mmas(perm);
      if (cnt < best) { best = cnt; for (int i = 0; i < na; ++i) best_perm[i] = perm[i]; }
    }
    if (best < d->num_taps) {
      for (int i = 0; i < na; ++i) k.acc_blk[i] = best_perm[i];
      // reorder the taps of every group: operands in order of first appearance, column blocks ascending inside an operand
      bool seen2[4] = {false, false, false, false};
      for (int g = 0; g < ng; ++g) {
        KGroup& G = hg[g];
        int order[kMaxGroupTaps], no = 0;
        bool done[kMaxGroupTaps] = {false};
        for (int j = 0; j < G.ntaps; ++j) {
          if (done[j]) continue;
          for (int c = 0; c < na; ++c)
            for (int q = j; q < G.ntaps; ++q)
              if (!done[q] && G.roff[q] == G.roff[j] && k.acc_blk[G.acc[q]] == c) { done[q] = true; order[no++] = q; }
        }
        KGroup R = G;
        for (int j = 0; j < G.ntaps; ++j) {
          const int q = order[j];
          R.roff[j] = G.roff[q]; R.acc[j] = G.acc[q]; R.brow[j] = G.brow[q];
        }
        G = R;
        for (int j = 0; j < G.ntaps; ++j) {
          G.first[j] = seen2[G.acc[j]] ? 0 : 1;
          seen2[G.acc[j]] = true;
        }
        // runs: consecutive taps, same operand, column blocks c, c+1, ..., the same accumulate-or-overwrite status
        for (int j = 0; j < G.ntaps;) {
          int len = 1;
          while (j + len < G.ntaps && len < max_run && G.roff[j + len] == G.roff[j] &&
                 k.acc_blk[G.acc[j + len]] == k.acc_blk[G.acc[j]] + len && G.first[j + len] == G.first[j])
            ++len;
          wm[g][j] = len;
          for (int q = 1; q < len; ++q) wm[g][j + q] = 0;
          j += len;
        }
      }
    }
  }
  {   // shared-memory slot of every tap's weight tile = its position in issue order (adjacent for merged taps)
    int t = 0;
    for (int g = 0; g < ng; ++g)
      for (int j = 0; j < hg[g].ntaps; ++j, ++t) hg[g].bidx[j] = t;
  }
  // ---- shared memory plan
  k.b_tap_bytes = ((d->block_n * k.row_bytes + 1023) / 1024) * 1024;
  const int b_total = halves * k.num_cblk * k.num_taps * k.b_tap_bytes;
  k.b_resident = (b_total <= resident_limit && halves * k.num_cblk <= kMaxStages) ? 1 : 0;
  for (int g = 0; g < ng; ++g)
    for (int j = 0; j < hg[g].ntaps; ++j) {
      hg[g].a16[j] = (hg[g].roff[j] * k.row_bytes) >> 4;
      hg[g].b16[j] = ((k.b_resident ? hg[g].bidx[j] : j) * k.b_tap_bytes) >> 4;
      hg[g].col[j] = (d->num_acc <= 4 ? k.acc_blk[hg[g].acc[j]] : hg[g].acc[j]) * d->block_n;
    }
  {
    int t = 0;
    for (int g = 0; g < ng; ++g)
      for (int j = 0; j < hg[g].ntaps; ++j, ++t) {
        k.f_a16[t] = hg[g].a16[j];
        k.f_b16[t] = hg[g].b16[j];
        k.f_col[t] = hg[g].col[j];
        k.f_flags[t] = (hg[g].first[j] ? 1 : 0) | (j == 0 ? 2 : 0) | (j == hg[g].ntaps - 1 ? 4 : 0);
        k.f_wm[t] = wm[g][j];
        if (wm[g][j] != 1) k.merged = 1;
      }
  }
  k.a_stage_bytes = (((k.TH + max_span) * k.TWB * k.row_bytes + 1023) / 1024) * 1024;
  k.b_stage_bytes = k.b_resident ? 0 : max_gt * k.b_tap_bytes;
  const int tiles_per_group = k.tiles_h * k.tiles_w;
  const int groups_total = d->n_img * k.n_blocks;
  const int sms = sfk_num_sms();
  int ctas_per_group = (2 * sms) / groups_total;
  if (ctas_per_group < 1) ctas_per_group = 1;
  if (ctas_per_group > tiles_per_group) ctas_per_group = tiles_per_group;
  const int total_ctas = ctas_per_group * groups_total;
  const int cols = d->num_acc * d->block_n * (k.m2 + 1);
  int per_sm = (total_ctas > sms && cols <= 256 && !k.m2) ? 2 : 1;
  const int stage_bytes = k.a_stage_bytes + k.b_stage_bytes;
  const int resident = k.b_resident ? b_total : 0;
  // staged TMA-store epilogue: whenever the tile's columns split into 64- (or 32-) column slabs and nothing is accumulated
  static const int ts_env = env_int("SFK_TMA_STORE", 0);   // 0 never (default), 1 wherever possible, 2 only 128-column tiles
  // Measured: the transpose pays where the epilogue is light (plain data gradients: 1024^2 332 -> 310 us) or the scatter is
  // widest (depth-to-space with Cq >= 64: 341 -> 317 us); the noise/bias/activation epilogues of narrow tiles are issue-bound
  // and lose to the 16 extra shuffles (1024^2 forward 395 -> 422 us, 512^2 202 -> 243 us, 1024^2 fused upsample 363 -> 395 us).
  static const int xs_env = env_int("SFK_XSTORE", 1);   // 0 never, 1 policy, 2 everywhere
  const int fl = d->flags & ~SFK_EP_PROFILE;
  k.xs = (!P->f32 && (xs_env == 2 || (xs_env == 1 && (fl == 0 || (d->out_d2s && d->out_c >= 256))))) ? 1 : 0;
  k.ts = 0;
  k.ts_slabw = 64;
  k.ts_nbuf = 2;
  if (!P->f32 && ts_env && !k.m2 && !(d->flags & SFK_EP_ACCUM) && d->block_n % 32 == 0) {
    if (d->out_d2s) {
      if (((d->out_c / 2) % 64) == 0) k.ts = 1;                 // a slab never straddles the two row phases
    } else {
      k.ts = 1;
      k.ts_slabw = d->block_n % 64 == 0 ? 64 : 32;
    }
  }
  int staging = k.ts ? 4 * k.ts_nbuf * 32 * k.ts_slabw * 2 : 0;
  const int cap1 = 216 * 1024;   // one CTA per SM: 227 KB opt-in limit minus the kernel's static shared memory
  if (per_sm == 2 && (100 * 1024 - resident - 1024 - staging) / stage_bytes < 2) per_sm = 1;
  if (k.ts && per_sm == 1 && (cap1 - resident - 1024 - staging) / stage_bytes < 2) {   // keep two pipeline stages: single-buffer
    k.ts_nbuf = 1;
    staging /= 2;
  }
  // Measured (8 pairs): the staged store gains a little where a tile has 128 columns and the staging can be double-buffered
  // (256^2 128->128 conv 184 -> 179 us, 256^2 fused upsample 359 -> 333 us); narrow tiles lose to the extra fence/wait latency
  // on the epilogue's critical path (1024^2 32->32 385 -> 412 us, 512^2 64->64 207 -> 248 us).  Whole step: off 10.74 ms,
  // 128-column policy 10.76-10.83 ms, everywhere 11.09 ms -> off by default.
  if (k.ts && ts_env == 2 && !(d->block_n == 128 && k.ts_nbuf == 2)) {
    k.ts = 0;
    staging = 0;
    if (total_ctas > sms && cols <= 256 && (100 * 1024 - resident - 1024) / stage_bytes >= 2) per_sm = 2;
  }
  const int budget = (per_sm == 2 ? 100 * 1024 : cap1) - resident - 1024 - staging;
  int stages = d->stages > 0 ? d->stages : budget / stage_bytes;
  const int ksteps_per_cta = k.num_cblk * k.passes * ng * ((tiles_per_group + ctas_per_group - 1) / ctas_per_group);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > ksteps_per_cta) stages = ksteps_per_cta;
  SFK_REQUIRE(stages >= 1 && stages * stage_bytes + resident + staging + 1024 <= 220 * 1024, SFK_E_SHAPE, "igemm: tile does not fit shared memory");
  k.ts_off = stages * stage_bytes + resident;
  k.stages = stages;
  const int tmem_limit = per_sm == 2 ? 256 : 512;
  k.acc_stages = (2 * cols <= tmem_limit) ? 2 : 1;
  k.ksplit = 1;
  if (P->f32) {   // partial accumulators against the truncating fp32 accumulation of the tensor core: as many as TMEM holds, up to 8
    static const int ksplit_env = env_int("SFK_KSPLIT", 8);
    int p = ksplit_env < 1 ? 1 : ksplit_env;
    while (p > 1 && (p * cols > tmem_limit || p > k.num_cblk * k.passes || (p & (p - 1)))) --p;
    k.ksplit = p;
    k.acc_stages = (2 * p * cols <= tmem_limit) ? 2 : 1;
  }
  k.dual_issue = (k.acc_stages == 2 && stages >= 2 * k.num_cblk * k.passes * ng) ? 1 : 0;
  const int want = k.acc_stages * cols * k.ksplit;
  k.tmem_cols = want <= 32 ? 32 : want <= 64 ? 64 : want <= 128 ? 128 : want <= 256 ? 256 : 512;

  for (int sp = 0; sp <= max_span; ++sp)
    SFK_REQUIRE(encode_a_map(enc, &k.mapA[sp], d, a_base, halves, es, KC, k.TWB, k.TH + sp, swz) == 0, SFK_E_DRIVER,
                "igemm: cuTensorMapEncodeTiled(A) failed");
  for (int sp = max_span + 1; sp < 4; ++sp) k.mapA[sp] = k.mapA[0];
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->a_c, (cuuint64_t)d->b_rows, (cuuint64_t)d->b_samples * halves};
    cuuint64_t strides[2] = {(cuuint64_t)d->a_c * es, (cuuint64_t)d->b_rows * d->a_c * es};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)d->block_n, 1};
    cuuint32_t es1[3] = {1, 1, 1};
    CUresult r = enc(&k.mapB, dt, 3, const_cast<void*>(b_base), dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SFK_REQUIRE(r == CUDA_SUCCESS, SFK_E_DRIVER, "igemm: cuTensorMapEncodeTiled(B) failed");
  }
  if (k.ts) {
    const cuuint64_t oc = static_cast<cuuint64_t>(d->out_c), ow = static_cast<cuuint64_t>(d->out_w), oh = static_cast<cuuint64_t>(d->out_h);
    cuuint64_t dims[5], strides[4];
    if (d->out_d2s) {   // [n][2*oh][2*ow][Cq] viewed as {pixel pair (2Cq), W, row phase, H, N}
      const cuuint64_t cq = oc / 4, frow = 2 * ow * cq * 2;
      dims[0] = 2 * cq; dims[1] = ow; dims[2] = 2; dims[3] = oh; dims[4] = static_cast<cuuint64_t>(d->n_img);
      strides[0] = 2 * cq * 2; strides[1] = frow; strides[2] = 2 * frow; strides[3] = 2 * oh * frow;
    } else {            // [n*num_acc][oh][ow][oc] with a unit dummy dimension in the row-phase slot
      dims[0] = oc; dims[1] = ow; dims[2] = 1; dims[3] = oh; dims[4] = static_cast<cuuint64_t>(d->n_img) * d->num_acc;
      strides[0] = oc * 2; strides[1] = ow * oc * 2; strides[2] = ow * oc * 2; strides[3] = oh * ow * oc * 2;
    }
    cuuint32_t box[5] = {static_cast<cuuint32_t>(k.ts_slabw), static_cast<cuuint32_t>(k.TW), 1, static_cast<cuuint32_t>(32 / k.TWB), 1};
    cuuint32_t es1[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&k.mapO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d->out, dims, strides, box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     k.ts_slabw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SFK_REQUIRE(r == CUDA_SUCCESS, SFK_E_DRIVER, "igemm: cuTensorMapEncodeTiled(out) failed");
  }
  for (int g = 0; g < ng; ++g) {
    KGroupDev& D = k.groups[g];
    D.plane = hg[g].plane; D.dx_min = hg[g].dx_min; D.dy_min = hg[g].dy_min; D.map = hg[g].map; D.bytes = hg[g].bytes; D.ntaps = hg[g].ntaps;
    for (int j = 0; j < kMaxGroupTaps; ++j) { D.brow[j] = hg[g].brow[j]; D.bidx[j] = hg[g].bidx[j]; }
  }
  static const int early_env = env_int("SFK_EARLY_RELEASE", 1);
  k.early = early_env ? 1 : 0;
  static_assert(sizeof(Igemm2Args) <= 4096, "kernel argument block must stay in the 4 KB constant bank window");
  P->smem = static_cast<size_t>(stages) * stage_bytes + resident + staging + 1024;
  P->grid = dim3(static_cast<unsigned>(ctas_per_group), static_cast<unsigned>(groups_total));
  P->fct = d->flags & ~SFK_EP_PROFILE;
  static const int spec_env = env_int("SFK_SPECIALIZE", 1);   // 0: always the run-time-variant kernels (A/B timing)
  P->var = (k.ts || !spec_env || (d->flags & SFK_EP_PROFILE)) ? -1 : ((k.out_d2s ? kVarD2S : 0) | (k.m2 ? kVarM2 : 0) | (k.xs ? kVarXS : 0) | (k.merged ? kVarMG : 0));
  // opt in to > 48 KB of dynamic shared memory: per (kernel instance, device), idempotent
  const void* fn = P->f32 ? kernel_for<float>(P->fct, P->var) : kernel_for<__nv_bfloat16>(P->fct, P->var);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) return static_cast<int>(e);
  return 0;
}

template <typename T>
int launch_ref(const sfk_igemm_desc* d, sfk_stream_t stream) {
  RefArgs<T> r;
  memset(&r, 0, sizeof(r));
  r.A = static_cast<const T*>(d->a);
  r.B = static_cast<const T*>(d->b);
  r.n_img = d->n_img; r.a_h = d->a_h; r.a_w = d->a_w; r.a_c = d->a_c; r.a_planes = d->a_planes;
  r.b_rows = d->b_rows; r.b_per_sample = d->b_samples > 1 ? 1 : 0;
  r.out_h = d->out_h; r.out_w = d->out_w; r.out_c = d->out_c; r.num_acc = d->num_acc; r.block_n = d->block_n;
  r.num_taps = d->num_taps; r.flags = d->flags; r.noise_w = d->noise_w;
  r.vec_stride = d->vec_stride > 0 ? d->vec_stride : d->out_c;
  r.out_d2s = d->out_d2s;
  r.a_s2d = d->a_s2d;
  r.out = static_cast<T*>(d->out);
  r.dscale = d->dscale; r.bias = d->bias; r.noise = d->noise;
  r.xin = static_cast<const T*>(d->xin); r.colscale = d->colscale; r.gs = d->gs;
  fill_taps(d, r.taps);
  const long total = static_cast<long>(d->n_img) * d->num_acc * d->out_h * d->out_w * (d->out_c / 16);
  long blocks = (total + 127) / 128;
  if (blocks > 65535L * 16) blocks = 65535L * 16;
  igemm_ref_kernel<T><<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(r);
  return sfk_check_launch("igemm_ref_kernel");
}


int launch_plan(const IgemmPlan* P, sfk_stream_t stream) {
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (P->use_ref) return launch_ref<float>(&P->desc, stream);
  if (P->k.passes == 3) {
    // operand copies for the split-tf32 passes (same stream: ordered before the GEMM)
    const int tpb = 256;
    long ba = (P->a_n4 + tpb - 1) / tpb, bb = (P->b_n4 + tpb - 1) / tpb;
    if (ba > 148 * 16) ba = 148 * 16;
    if (bb > 148 * 16) bb = 148 * 16;
    tf32_split_kernel<<<static_cast<unsigned>(ba), tpb, 0, cs>>>(reinterpret_cast<const float4*>(P->a_src), reinterpret_cast<float4*>(P->a_hi),
                                                                reinterpret_cast<float4*>(P->a_hi) + P->a_n4, P->a_n4);
    tf32_split_kernel<<<static_cast<unsigned>(bb), tpb, 0, cs>>>(reinterpret_cast<const float4*>(P->b_src), reinterpret_cast<float4*>(P->b_hi),
                                                                reinterpret_cast<float4*>(P->b_hi) + P->b_n4, P->b_n4);
  }
  const void* fn = P->f32 ? kernel_for<float>(P->fct, P->var) : kernel_for<__nv_bfloat16>(P->fct, P->var);
  void* args[1] = {const_cast<Igemm2Args*>(&P->k)};
  cudaError_t e = cudaLaunchKernel(fn, P->grid, dim3(kThreads), args, P->smem, cs);
  if (e != cudaSuccess) {
    sfk_set_error(cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return sfk_check_launch("igemm_tc2_kernel");
}
}  // namespace

extern "C" int sfk_set_conv_math(int mode) {
  if (mode < 0 || mode > 3) return SFK_E_ARG;
  g_conv_math = mode;
  return 0;
}
extern "C" int sfk_get_conv_math(void) { return g_conv_math; }

extern "C" size_t sfk_igemm_workspace_bytes(const sfk_igemm_desc* d) {
  if (!d || !sfk_act_f32() || g_conv_math == 1 || g_conv_math == 3) return 0;
  return split_ws_bytes(d);
}

// One-shot form: plans on the caller's stack frame (heap for the 5 KB argument block), launches, forgets.  Re-entrant.
extern "C" int sfk_igemm(const sfk_igemm_desc* d, sfk_stream_t stream) {
  IgemmPlan* P = static_cast<IgemmPlan*>(malloc(sizeof(IgemmPlan)));
  if (!P) return SFK_E_ARG;
  int rc = build_plan(d, P);
  if (rc == 0) rc = launch_plan(P, stream);
  free(P);
  return rc;
}

// Prepared form: everything host-side (tap grouping, shared-memory plan, tensor-map encoding) happens once; sfk_igemm_run is a
// single cudaLaunchKernel.  A plan is immutable, so any number of host threads / streams may run it concurrently.
extern "C" int sfk_igemm_prepare(const sfk_igemm_desc* d, sfk_igemm_plan** out) {
  SFK_REQUIRE(out != nullptr, SFK_E_ARG, "igemm_prepare: null output");
  IgemmPlan* P = static_cast<IgemmPlan*>(malloc(sizeof(IgemmPlan)));
  SFK_REQUIRE(P != nullptr, SFK_E_ARG, "igemm_prepare: out of host memory");
  const int rc = build_plan(d, P);
  if (rc) {
    free(P);
    return rc;
  }
  *out = reinterpret_cast<sfk_igemm_plan*>(P);
  return 0;
}
extern "C" int sfk_igemm_run(const sfk_igemm_plan* plan, sfk_stream_t stream) {
  SFK_REQUIRE(plan != nullptr, SFK_E_ARG, "igemm_run: null plan");
  return launch_plan(reinterpret_cast<const IgemmPlan*>(plan), stream);
}
// What the planner decided, for tests and for the launch <-> ncu-row map under profiles/:
//   [0] two M tiles per stage  [1] halo loads  [2] resident weights  [3] depth-to-space out  [4] space-to-depth in  [5] passes
//   [6] pipeline stages  [7] block_n  [8] compile-time variant (-1 run time)  [9] epilogue flags  [10] grid.x  [11] grid.y
//   [12] dynamic smem bytes  [13] fp32 storage  [14] CUDA-core kernel  [15] accumulator stages * 16 + partial accumulators
extern "C" int sfk_igemm_plan_info(const sfk_igemm_plan* plan, int32_t* out16) {
  SFK_REQUIRE(plan != nullptr && out16 != nullptr, SFK_E_ARG, "igemm_plan_info: null");
  const IgemmPlan* P = reinterpret_cast<const IgemmPlan*>(plan);
  const Igemm2Args& k = P->k;
  const int v[16] = {k.m2, k.TWB != k.TW ? 1 : 0, k.b_resident, k.out_d2s, k.a_s2d, k.passes, k.stages, k.block_n, P->var, P->fct,
                     static_cast<int>(P->grid.x), static_cast<int>(P->grid.y), static_cast<int>(P->smem), P->f32, P->use_ref, k.acc_stages * 16 + k.ksplit};
  for (int i = 0; i < 16; ++i) out16[i] = v[i];
  return 0;
}
extern "C" int sfk_igemm_destroy(sfk_igemm_plan* plan) {
  free(plan);
  return 0;
}


extern "C" int sfk_role_cycles(unsigned long long* out8, int reset) {
  if (out8 && cudaMemcpyFromSymbol(out8, g_role_cycles, sizeof(unsigned long long) * 8) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(g_role_cycles, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}

extern "C" int sfk_igemm_ref(const sfk_igemm_desc* d, sfk_stream_t stream) {
  int rc = validate(d);
  if (rc) return rc;
  return sfk_act_f32() ? launch_ref<float>(d, stream) : launch_ref<__nv_bfloat16>(d, stream);
}
