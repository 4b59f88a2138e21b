// Bandwidth-bound kernels of the attack path: 128-bit vectorised, channel-innermost (NHWC) so that a
// warp touches contiguous bytes, per-channel reductions done warp/block-first and flushed with a
// handful of global atomics per block.  Grids are sized in multiples of the SM count.
#include "sfk_common.cuh"

namespace {

constexpr int kBlock = 256;
using bf16 = __nv_bfloat16;

inline unsigned grid_for(long work_items, int per_sm_blocks = 8) {
  long blocks = (work_items + kBlock - 1) / kBlock;
  long cap = static_cast<long>(sfk_num_sms()) * per_sm_blocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

// flush per-thread channel-vector accumulators (thread owns channels [cv*8, cv*8+8)) to global[n][C]
__device__ __forceinline__ void flush_channel_acc(float* sacc, const float* acc, int cv, int C, float* gdst) {
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&sacc[cv * 8 + i], acc[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(gdst + i, sacc[i]);
}

// ============================================================================================
// first-layer conv (Cin = 3).  CUDA-core kernels, register-blocked so that they are bound by FMA issue and not by the
// shared-memory weight fetches: one thread owns 4 horizontally adjacent pixels x 8 output channels, so every weight vector
// read from shared memory feeds 32 FMAs (forward) / 96 FMAs (backward).
template <typename T>
__global__ void __launch_bounds__(kBlock) conv_c3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                             T* __restrict__ out, int N, int H, int W, int Cout, int relu) {
  extern __shared__ __align__(16) float sw[];  // [27][Cout] then bias[Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;  // k = c*9 + ky*3 + kx
    sw[i] = w[co * 27 + k];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int groups = Cout / 8, W4 = (W + 3) / 4;
  const long total = static_cast<long>(N) * H * W4 * groups;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = idx % groups;
    long p = idx / groups;
    const int w0 = (p % W4) * 4;
    p /= W4;
    const int hq = p % H;
    const int n = p / H;
    float acc[4][8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(sw + 27 * Cout + g * 8), b1 = *reinterpret_cast<const float4*>(sw + 27 * Cout + g * 8 + 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
        acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* xp = x + (static_cast<long>(n) * 3 + c) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int ih = hq + ky - 1;
        if (ih < 0 || ih >= H) continue;
        float xr[6];   // input columns w0-1 .. w0+4 of this row
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int iw = w0 + q - 1;
          xr[q] = (iw >= 0 && iw < W) ? __ldg(xp + static_cast<long>(ih) * W + iw) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wp = sw + (c * 9 + ky * 3 + kx) * Cout + g * 8;
          const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(xr[j + kx], wv[i], acc[j][i]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (w0 + j >= W) break;
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaxf(acc[j][i], 0.f);
      }
      store8(out + ((static_cast<long>(n) * H + hq) * W + w0 + j) * Cout + g * 8, acc[j]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) conv_c3_bwd_kernel(const T* __restrict__ g, const float* __restrict__ w, float* __restrict__ gx, int N,
                                                             int H, int W, int Cout, int LP /* lanes per pixel quad, power of 2 <= 8 */) {
  // [9][Cout/8][8*4 + 4]: input channel padded to 4 (one LDS.128 per output channel); each 8-channel vector padded by 4 floats so
  // that the lanes of a pixel quad (one vector each) read from distinct banks
  extern __shared__ __align__(16) float sw[];
  const int vecs = Cout / 8, W4 = (W + 3) / 4;
  for (int i = threadIdx.x; i < 36 * Cout; i += blockDim.x) {
    const int c = i % 4, co = (i / 4) % Cout, k = i / (4 * Cout);
    sw[(k * vecs + co / 8) * 36 + (co % 8) * 4 + c] = c < 3 ? w[co * 27 + c * 9 + k] : 0.f;
  }
  __syncthreads();
  const long total = static_cast<long>(N) * H * W4 * LP;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  const long rounds = (total + stride - 1) / stride;
  for (long rr = 0; rr < rounds; ++rr) {
    const long idx = rr * stride + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const bool active = idx < total;
    const int sub = idx % LP;
    long p = idx / LP;
    const int w0 = (p % W4) * 4;
    p /= W4;
    const int hq = p % H;
    const int n = p / H;
    float a[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j][0] = a[j][1] = a[j][2] = 0.f;
    if (active) {
      for (int ky = 0; ky < 3; ++ky) {
        const int oh = hq - ky + 1;
        if (oh < 0 || oh >= H) continue;
        const T* grow = g + (static_cast<long>(n) * H + oh) * W * Cout;
        for (int v = sub; v < vecs; v += LP) {
          // the six gradient vectors this row contributes to the four pixels (columns w0-1 .. w0+4), all loads in flight at once
          float gc[6][8];
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const int ow = w0 + c - 1;
            if (ow >= 0 && ow < W) {
              load8(grow + static_cast<long>(ow) * Cout + v * 8, gc[c]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) gc[c][i] = 0.f;
            }
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * vecs + v) * 36);
            float4 wv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) wv[i] = wp[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float* gv = gc[j - kx + 2];      // output column ow = w0 + j - kx + 1
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                a[j][0] = fmaf(gv[i], wv[i].x, a[j][0]);
                a[j][1] = fmaf(gv[i], wv[i].y, a[j][1]);
                a[j][2] = fmaf(gv[i], wv[i].z, a[j][2]);
              }
            }
          }
        }
      }
    }
    for (int o = LP >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j][0] += __shfl_xor_sync(0xffffffffu, a[j][0], o);
        a[j][1] += __shfl_xor_sync(0xffffffffu, a[j][1], o);
        a[j][2] += __shfl_xor_sync(0xffffffffu, a[j][2], o);
      }
    }
    if (active && sub == 0) {
      const long hw = static_cast<long>(H) * W;
      float* o = gx + static_cast<long>(n) * 3 * hw + static_cast<long>(hq) * W + w0;
      if ((W & 3) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(o + c * hw) = make_float4(a[0][c], a[1][c], a[2][c], a[3][c]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (w0 + j < W) {
            o[j] = a[j][0];
            o[hw + j] = a[j][1];
            o[2 * hw + j] = a[j][2];
          }
        }
      }
    }
  }
}

// ============================================================================================
// 3-channel image <-> 16-channel NHWC bf16 operand of the tensor-core conv (first conv of VGG / the encoder on tcgen05).
// A bf16 copy of the image would lose the perturbation (2/255 steps against 2^-8 relative precision), so the image is split
// x = hi + lo (both bf16, error 2^-17 relative) and the weights likewise; the 16 K-channels carry
//   operand  [x.hi(3) | x.lo(3) | x.hi(3) | 0 x 7]     weights  [W.hi(3) | W.hi(3) | W.lo(3) | 0 x 7]
// so that one K = 16 MMA per tap accumulates x.hi*W.hi + x.lo*W.hi + x.hi*W.lo in fp32 (the lo*lo term, 2^-18, is dropped).
__global__ void c3_pack_kernel(const float* __restrict__ x, uint4* __restrict__ xp, int N, int HW) {
  const long total = static_cast<long>(N) * HW;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = idx / HW, p = idx - n * HW;
    const float* xb = x + n * 3 * HW + p;
    float hi[3], lo[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldg(xb + static_cast<long>(c) * HW);
      hi[c] = __bfloat162float(__float2bfloat16(v));
      lo[c] = v - hi[c];
    }
    uint4 a, b;
    a.x = pack2(hi[0], hi[1]);
    a.y = pack2(hi[2], lo[0]);
    a.z = pack2(lo[1], lo[2]);
    a.w = pack2(hi[0], hi[1]);
    b.x = pack2(hi[2], 0.f);
    b.y = b.z = b.w = 0u;
    xp[2 * idx] = a;
    xp[2 * idx + 1] = b;
  }
}

// gradient of the same operand back to the image: gx[n][c][p] = gp[n][p][c] + gp[n][p][3 + c]  (the W.hi and W.lo rows of the
// transposed weights; fp32 sum of the two bf16 parts)
__global__ void c3_unpack_kernel(const uint4* __restrict__ gp, float* __restrict__ gx, int N, int HW) {
  const long total = static_cast<long>(N) * HW;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = idx / HW, p = idx - n * HW;
    float f[8];
    unpack8(__ldg(gp + 2 * idx), f);
    float* gb = gx + n * 3 * HW + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) gb[static_cast<long>(c) * HW] = f[c] + f[3 + c];
  }
}

// ============================================================================================
// pools
__global__ void avgpool_affine_kernel(const float* __restrict__ x, float* __restrict__ y, long planes, int H, int W, int k,
                                      float a, float b) {
  const int Ho = H / k, Wo = W / k;
  const long total = planes * Ho * Wo;
  const float inv = a / static_cast<float>(k * k);
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int j = idx % Wo;
    long p = idx / Wo;
    const int i = p % Ho;
    const long pl = p / Ho;
    const float* xp = x + (pl * H + static_cast<long>(i) * k) * W + static_cast<long>(j) * k;
    float s = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) s += __ldg(xp + static_cast<long>(dy) * W + dx);
    y[idx] = s * inv + b;
  }
}

template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, vecs = C / 8;
  const long total = static_cast<long>(N) * Ho * Wo * vecs;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    long p = idx / vecs;
    const int j = p % Wo;
    p /= Wo;
    const int i = p % Ho;
    const int n = p / Ho;
    float m[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) m[q] = -INFINITY;
    for (int dy = 0; dy < 2; ++dy) {
      const int h = 2 * i + dy;
      if (h >= H) continue;
      for (int dx = 0; dx < 2; ++dx) {
        const int w = 2 * j + dx;
        if (w >= W) continue;
        float t[8];
        load8(x + ((static_cast<long>(n) * H + h) * W + w) * C + v * 8, t);
#pragma unroll
        for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], t[q]);
      }
    }
    store8(y + ((static_cast<long>(n) * Ho + i) * Wo + j) * C + v * 8, m);
  }
}

// One thread per pooling window x 8 channels: routes gy to the FIRST maximum in row-major window order
// (torch's tie-break; ties are common in bf16), adds the optional feature-tap gradient, applies the ReLU mask.
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, T* __restrict__ gx,
                                    const T* __restrict__ tap_ref, float tap_coef, int relu_mask, int N, int H, int W, int C) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, vecs = C / 8;
  const long total = static_cast<long>(N) * Ho * Wo * vecs;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    long p = idx / vecs;
    const int j = p % Wo;
    p /= Wo;
    const int i = p % Ho;
    const int n = p / Ho;
    float xin[4][8], rin[4][8];
    bool inb[4];
    float m[8], g[8];
    int am[8];
    load8(gy + ((static_cast<long>(n) * Ho + i) * Wo + j) * C + v * 8, g);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      m[q] = -INFINITY;
      am[q] = -1;
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int h = 2 * i + (s >> 1), w = 2 * j + (s & 1);
      inb[s] = (h < H) && (w < W);
      if (inb[s]) {
        load8(x + ((static_cast<long>(n) * H + h) * W + w) * C + v * 8, xin[s]);
        if (tap_ref) load8(tap_ref + ((static_cast<long>(n) * H + h) * W + w) * C + v * 8, rin[s]);   // all loads up front
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (xin[s][q] > m[q]) {
            m[q] = xin[s][q];
            am[q] = s;
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (!inb[s]) continue;
      const int h = 2 * i + (s >> 1), w = 2 * j + (s & 1);
      const long off = ((static_cast<long>(n) * H + h) * W + w) * C + v * 8;
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = (am[q] == s) ? g[q] : 0.f;
      if (tap_ref) {
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] += tap_coef * (xin[s][q] - rin[s][q]);
      }
      if (relu_mask) {
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = xin[s][q] > 0.f ? o[q] : 0.f;
      }
      store8(gx + off, o);
    }
  }
}

template <typename T>
__global__ void gap_fwd_kernel(const T* __restrict__ x, float* __restrict__ y, int HW, int C) {
  // block = (n, channel-vector group); threads stride over HW
  const int n = blockIdx.y, vecs = C / 8;
  const int v = blockIdx.x;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float t[8];
    load8(x + (static_cast<long>(n) * HW + p) * C + v * 8, t);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += t[q];
  }
  __shared__ float red[8][kBlock / 32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float s = warp_sum(acc[q]);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int k = 0; k < blockDim.x / 32; ++k) s += red[threadIdx.x][k];
    y[static_cast<long>(n) * C + v * 8 + threadIdx.x] = s / static_cast<float>(HW);
  }
  (void)vecs;
}

template <typename T>
__global__ void gap_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gy, T* __restrict__ gx, int N, int HW, int C) {
  const int vecs = C / 8;
  const long total = static_cast<long>(N) * HW * vecs;
  const float inv = 1.f / static_cast<float>(HW);
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    const long p = idx / vecs;
    const int n = p / HW;
    float t[8], o[8];
    load8(x + p * C + v * 8, t);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = t[q] > 0.f ? gy[static_cast<long>(n) * C + v * 8 + q] * inv : 0.f;
    store8(gx + p * C + v * 8, o);
  }
}

// ============================================================================================
// losses
template <typename T>
__global__ void mse_tap_kernel(const T* __restrict__ f, const T* __restrict__ ref, T* __restrict__ g, float* __restrict__ loss,
                               float coef_loss, float coef_grad, int accumulate, int relu_mask, long per_sample) {
  const int n = blockIdx.y;
  const long vecs = per_sample / 8;
  const long base = static_cast<long>(n) * per_sample;
  float lsum = 0.f;
  for (long v = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; v < vecs; v += static_cast<long>(gridDim.x) * blockDim.x) {
    float a[8], r[8];
    load8(f + base + v * 8, a);
    load8(ref + base + v * 8, r);
    float o[8];
    if (g && accumulate) load8p(g + base + v * 8, o);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float d = a[q] - r[q];
      lsum += d * d;
      float gq = coef_grad * d;
      if (relu_mask && !(a[q] > 0.f)) gq = 0.f;
      o[q] = (g && accumulate) ? o[q] + gq : gq;
    }
    if (g) store8(g + base + v * 8, o);
  }
  lsum = warp_sum(lsum);
  __shared__ float red[kBlock / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float s = 0.f;
    for (int k = 0; k < blockDim.x / 32; ++k) s += red[k];
    atomicAdd(loss + n, coef_loss * s);
  }
}

// fp32 variant for latent codes: loss[n] += coef_loss*sum((a-b)^2); g (=|+=) coef_grad*(a-b)
__global__ void mse_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ g, float* __restrict__ loss,
                               float coef_loss, float coef_grad, int accumulate, long per_sample) {
  const int n = blockIdx.y;
  const long base = static_cast<long>(n) * per_sample;
  float lsum = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per_sample; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float d = a[base + i] - b[base + i];
    lsum += d * d;
    if (g) g[base + i] = (accumulate ? g[base + i] : 0.f) + coef_grad * d;
  }
  lsum = warp_sum(lsum);
  __shared__ float red[kBlock / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float t = 0.f;
    for (int q = 0; q < blockDim.x / 32; ++q) t += red[q];
    atomicAdd(loss + n, coef_loss * t);
  }
}

__global__ void image_loss_grad_kernel(const float* __restrict__ img, const float* __restrict__ ref, const float* __restrict__ gpool,
                                       float* __restrict__ g, float* __restrict__ loss, float coef_loss, float coef_grad, int S, int k) {
  const int n = blockIdx.y;
  const long per4 = 3L * S * S / 4;
  const int Sp = S / k, S4 = S / 4;
  const float invk2 = 1.f / static_cast<float>(k * k);
  float lsum = 0.f;
  const float4* iv = reinterpret_cast<const float4*>(img + static_cast<long>(n) * 3 * S * S);
  const float4* rv = reinterpret_cast<const float4*>(ref + static_cast<long>(n) * 3 * S * S);
  float4* gv = reinterpret_cast<float4*>(g + static_cast<long>(n) * 3 * S * S);
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int w4 = i % S4;
    long p = i / S4;
    const int h = p % S;
    const int c = p / S;
    const float4 a = __ldg(iv + i), r = __ldg(rv + i);
    const float* ap = reinterpret_cast<const float*>(&a);
    const float* rp = reinterpret_cast<const float*>(&r);
    float4 o;
    float* op = reinterpret_cast<float*>(&o);
    const float* gr = gpool ? gpool + ((static_cast<long>(n) * 3 + c) * Sp + h / k) * Sp : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = ap[j] - rp[j];
      lsum += d * d;
      op[j] = coef_grad * d + (gr ? invk2 * __ldg(gr + (w4 * 4 + j) / k) : 0.f);
    }
    gv[i] = o;
  }
  lsum = warp_sum(lsum);
  __shared__ float red[kBlock / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float t = 0.f;
    for (int q = 0; q < blockDim.x / 32; ++q) t += red[q];
    atomicAdd(loss + n, coef_loss * t);
  }
}

// ============================================================================================
// style space
__global__ void style_affine_fwd_kernel(const float* __restrict__ w, const float* __restrict__ A, const float* __restrict__ bias,
                                        const int* __restrict__ row_widx, float* __restrict__ s, int N, int L, int D, int SD, float scale) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < SD; r += warps) {
    const int l = row_widx[r];
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = lane; k < D; k += 32) acc = fmaf(__ldg(A + static_cast<long>(r) * D + k), __ldg(w + (static_cast<long>(n) * L + l) * D + k), acc);
      acc = warp_sum(acc);
      if (lane == 0) s[static_cast<long>(n) * SD + r] = bias[r] + scale * acc;
    }
  }
}

// gw[n][l][k] = scale * sum over the rows r of every layer driven by latent l of gs[n][r] * A[r][k].
// block = (layer, slice of its rows, group of 8 samples): A is streamed once per sample group, coalesced along k; partial sums
// are merged with atomics (gw is zeroed by the launcher).
constexpr int kAffRows = 32;
__global__ void style_affine_bwd_kernel(const float* __restrict__ gs, const float* __restrict__ A, const int* __restrict__ layer_row_start,
                                        const int* __restrict__ layer_widx, float* __restrict__ gw, int N, int L, int D, int SD, float scale) {
  __shared__ float sg[8][kAffRows];
  const int ly = blockIdx.x;
  const int l = layer_widx[ly];
  const int n0 = blockIdx.z * 8;
  const int rend = layer_row_start[ly + 1];
  for (int r0 = layer_row_start[ly] + blockIdx.y * kAffRows; r0 < rend; r0 += gridDim.y * kAffRows) {
  const int r1 = min(rend, r0 + kAffRows);
  __syncthreads();
  for (int t = threadIdx.x; t < 8 * kAffRows; t += blockDim.x) {
    const int q = t / kAffRows, r = r0 + t % kAffRows;
    sg[q][t % kAffRows] = (n0 + q < N && r < r1) ? __ldg(gs + static_cast<long>(n0 + q) * SD + r) : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float a = __ldg(A + static_cast<long>(r) * D + k);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = fmaf(sg[q][r - r0], a, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (n0 + q < N) atomicAdd(gw + (static_cast<long>(n0 + q) * L + l) * D + k, scale * acc[q]);
  }
  }
}

__global__ void demod_fwd_kernel(const float* __restrict__ s, int s_stride, const float* __restrict__ Q, float* __restrict__ d, int N, int Cin, int Cout) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = N * Cout;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total; r += warps) {
    const int j = r % Cout, n = r / Cout;
    float acc = 0.f;
    for (int i = lane; i < Cin; i += 32) {
      const float sv = __ldg(s + static_cast<long>(n) * s_stride + i);
      acc = fmaf(sv * sv, __ldg(Q + static_cast<long>(j) * Cin + i), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) d[r] = rsqrtf(acc + 1e-8f);
  }
}

// block = 32 input channels x 8 slices of the output-channel sum; Q rows are read coalesced along i
__global__ void demod_bwd_kernel(const float* __restrict__ s, int s_stride, const float* __restrict__ Q, const float* __restrict__ d,
                                 const float* __restrict__ gdacc, float* __restrict__ gs, int gs_stride, int N, int Cin, int Cout) {
  __shared__ float red[8][33];
  const int n = blockIdx.y;
  const int li = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + li;
  float acc = 0.f;
  if (i < Cin) {
    for (int j = sl; j < Cout; j += 8) {
      const float dj = __ldg(d + n * Cout + j);
      acc = fmaf(__ldg(gdacc + n * Cout + j) * dj * dj, __ldg(Q + static_cast<long>(j) * Cin + i), acc);
    }
  }
  red[sl][li] = acc;
  __syncthreads();
  if (sl == 0 && i < Cin) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][li];
    gs[static_cast<long>(n) * gs_stride + i] -= __ldg(s + static_cast<long>(n) * s_stride + i) * t;
  }
  (void)N;
}

template <typename T>
__global__ void modulate_weights_kernel(const float* __restrict__ wbase, const float* __restrict__ s, int s_stride, T* __restrict__ wmod,
                                        int N, long rows /* taps*cout */, int Cin, const float* __restrict__ d, int cout, int d_cols) {
  const int vecs = Cin / 8;
  const long total = static_cast<long>(N) * rows * vecs;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    const long r = (idx / vecs) % rows;
    const int n = idx / (vecs * rows);
    const float4* wp = reinterpret_cast<const float4*>(wbase + r * Cin + v * 8);
    const float4* sp = reinterpret_cast<const float4*>(s + static_cast<long>(n) * s_stride + v * 8);
    const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
    float o[8] = {w0.x * s0.x, w0.y * s0.y, w0.z * s0.z, w0.w * s0.w, w1.x * s1.x, w1.y * s1.y, w1.z * s1.z, w1.w * s1.w};
    if (d != nullptr) {   // demodulation folded into the weights: the conv epilogue then has no per-column scale to fetch
      const float dv = __ldg(d + static_cast<long>(n) * d_cols + (r % cout) % d_cols);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] *= dv;
    }
    store8(wmod + (static_cast<long>(n) * rows + r) * Cin + v * 8, o);
  }
}

// ---------------------------------------------------------------------------------------------
// All modulated-conv layers of the generator in ONE launch each (the per-layer kernels above stay for single-layer callers):
// the styles of every layer are known before the first conv runs, and the demodulation gradient is only consumed after the last.
// Layer table (device, int64 [n_layers][SFK_STYLE_TAB_COLS]): s_off, cin, cout, q_off, d_off, rows, wb_off, wm_off, d_cols, fold
//   q_off / d_off / wb_off / wm_off index the concatenated Q (cout x cin), d and gdacc ([n][cout] per layer), base weights
//   ([rows][cin]) and modulated weights ([n][rows][cin] per layer) buffers; rows = taps * (cout or 4*cout); fold != 0 folds d in,
//   fold == 2 also the activation gain sqrt(2) (for the SFK_EP_LRELU_RAW epilogue).
__global__ void demod_fwd_batched_kernel(const float* __restrict__ s, int s_stride, const float* __restrict__ Qc, float* __restrict__ dc,
                                         const long long* __restrict__ tab, int N) {
  const long long* t = tab + static_cast<long>(blockIdx.y) * SFK_STYLE_TAB_COLS;
  const int s_off = static_cast<int>(t[0]), Cin = static_cast<int>(t[1]), Cout = static_cast<int>(t[2]);
  const float* Q = Qc + t[3];
  float* d = dc + t[4];
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = N * Cout;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total; r += warps) {
    const int j = r % Cout, n = r / Cout;
    float acc = 0.f;
    for (int i = lane; i < Cin; i += 32) {
      const float sv = __ldg(s + static_cast<long>(n) * s_stride + s_off + i);
      acc = fmaf(sv * sv, __ldg(Q + static_cast<long>(j) * Cin + i), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) d[r] = rsqrtf(acc + 1e-8f);
  }
}

template <typename T>
__global__ void modulate_weights_batched_kernel(const float* __restrict__ wbc, const float* __restrict__ s, int s_stride, T* __restrict__ wmc,
                                                const float* __restrict__ dc, const long long* __restrict__ tab, int N) {
  const long long* t = tab + static_cast<long>(blockIdx.y) * SFK_STYLE_TAB_COLS;
  const int s_off = static_cast<int>(t[0]), Cin = static_cast<int>(t[1]), Cout = static_cast<int>(t[2]);
  const long rows = t[5];
  const float* wbase = wbc + t[6];
  T* wmod = wmc + t[7] * N;
  const int d_cols = static_cast<int>(t[8]);
  const float* d = t[9] ? dc + t[4] : nullptr;
  const float gain = t[9] == 2 ? SFK_SQRT2 : 1.f;
  const int rows_per_tap = static_cast<int>(rows / 9);
  const int vecs = Cin / 8;
  const long total = rows * vecs;
  // one thread = one 8-wide weight vector for EVERY sample: the fp32 base weights are read once, not once per sample
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    const long r = idx / vecs;
    const float4* wp = reinterpret_cast<const float4*>(wbase + r * Cin + v * 8);
    const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
    const int dcol = (r % rows_per_tap) % d_cols;
    for (int n = 0; n < N; ++n) {
      const float4* sp = reinterpret_cast<const float4*>(s + static_cast<long>(n) * s_stride + s_off + v * 8);
      const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
      float o[8] = {w0.x * s0.x, w0.y * s0.y, w0.z * s0.z, w0.w * s0.w, w1.x * s1.x, w1.y * s1.y, w1.z * s1.z, w1.w * s1.w};
      if (d != nullptr) {
        const float dv = __ldg(d + static_cast<long>(n) * d_cols + dcol) * gain;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] *= dv;
      }
      store8(wmod + (static_cast<long>(n) * rows + r) * Cin + v * 8, o);
    }
  }
  (void)Cout;
}

__global__ void demod_bwd_batched_kernel(const float* __restrict__ s, int s_stride, const float* __restrict__ Qc, const float* __restrict__ dc,
                                         const float* __restrict__ gdc, float* __restrict__ gs, int gs_stride, const long long* __restrict__ tab) {
  __shared__ float red[8][33];
  const long long* t = tab + static_cast<long>(blockIdx.z) * SFK_STYLE_TAB_COLS;
  const int s_off = static_cast<int>(t[0]), Cin = static_cast<int>(t[1]), Cout = static_cast<int>(t[2]);
  if (static_cast<int>(blockIdx.x) * 32 >= Cin) return;
  const float* Q = Qc + t[3];
  const float* d = dc + t[4];
  const float* gdacc = gdc + t[4];
  const int n = blockIdx.y;
  const int li = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + li;
  float acc = 0.f;
  if (i < Cin) {
    for (int j = sl; j < Cout; j += 8) {
      const float dj = __ldg(d + n * Cout + j);
      acc = fmaf(__ldg(gdacc + n * Cout + j) * dj * dj, __ldg(Q + static_cast<long>(j) * Cin + i), acc);
    }
  }
  red[sl][li] = acc;
  __syncthreads();
  if (sl == 0 && i < Cin) {
    float tt = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) tt += red[q][li];
    gs[static_cast<long>(n) * gs_stride + s_off + i] -= __ldg(s + static_cast<long>(n) * s_stride + s_off + i) * tt;
  }
}

// ============================================================================================
// blur (upfirdn2d [1,3,3,1], pad (1,1)) over the phase-planar transposed-conv output, fused with
// demod, noise, bias, leaky-relu
__device__ __forceinline__ float blur_w(int t) { return (t == 0 || t == 3) ? 0.25f : 0.75f; }

// Register-blocked separable 4x4 FIR: one thread produces a 2 x 4 block of positions x 8 channels.  Input rows are streamed:
// each of the 5 rows is loaded once (7 vectors), blurred horizontally into 4 partial columns, then scattered into the 2 output
// rows -- 35 loads for 8 outputs (4.4 per output instead of 16).
//   acc[i][j] = sum_{t,u} in[row0+i+t][col0+j+u] * k[t] * k[u]
template <class Load>
__device__ __forceinline__ void blur_block_2x4(int row0, int col0, Load&& ld, float (&acc)[2][4][8]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][j][c] = 0.f;
#pragma unroll
  for (int rr = 0; rr < 5; ++rr) {
    float h[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) h[j][c] = 0.f;
#pragma unroll
    for (int cc = 0; cc < 7; ++cc) {
      float v[8];
      ld(row0 + rr, col0 + cc, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = cc - j;
        if (u >= 0 && u < 4) {
          const float wu = blur_w(u);
#pragma unroll
          for (int c = 0; c < 8; ++c) h[j][c] = fmaf(wu, v[c], h[j][c]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int t = rr - i;
      if (t >= 0 && t < 4) {
        const float wt = blur_w(t);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][j][c] = fmaf(wt, h[j][c], acc[i][j][c]);
      }
    }
  }
}

// Shared-memory tiled version of the blur (forward) and its transpose (backward).
//   load phase   : the 19 x (TC+3) input halo of a 16 x TC output tile is fetched ONCE, 16 B per thread, all loads independent
//                  (backward: gy = slope(out)*gout is formed here and the demodulation reduction sum gy*y is taken on the
//                  tile's own 16 x TC pixels);
//   compute phase: each thread produces a 2 x 4 block of positions x 8 channels from shared memory with the separable
//                  register-blocked FIR above (35 LDS.128 per 8 outputs);
//   bank layout  : positions are 64 B (CV=4 channel vectors) apart with 64 B of padding after every 4th one, so the two
//                  column blocks served in one LDS.128 phase fall into different halves of the 32 banks.
// Straightforward (one thread = one position x 8 channels, 16 taps) versions of the blur and its transpose, generic in the
// storage type: used by the fp32 parity mode (sfk_set_activation_dtype(1)), where speed is irrelevant.
template <typename T>
__global__ void blur_simple_fwd_kernel(const T* __restrict__ Tn_, T* __restrict__ out, const float* __restrict__ d, const float* __restrict__ noise,
                                       float noise_w, const float* __restrict__ bias, int H, int W, int C) {
  const int n = blockIdx.y;
  const int vecs = C / 8, Ho = 2 * H, Wo = 2 * W, Hp = H + 1, Wp = W + 1;
  const long total = static_cast<long>(Ho) * Wo * vecs;
  const T* Tn = Tn_ + static_cast<long>(n) * 4 * Hp * Wp * C;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    const long p = idx / vecs;
    const int pw = p % Wo, po = p / Wo;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < 4; ++t) {
      const int q = po + t - 1;
      if (q < 0 || q > 2 * H) continue;
      for (int u = 0; u < 4; ++u) {
        const int r = pw + u - 1;
        if (r < 0 || r > 2 * W) continue;
        float tv[8];
        load8(Tn + ((static_cast<long>((q & 1) * 2 + (r & 1)) * Hp + (q >> 1)) * Wp + (r >> 1)) * C + v * 8, tv);
        const float wgt = blur_w(t) * blur_w(u);
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, tv[i], acc[i]);
      }
    }
    const float nz = noise ? noise_w * noise[static_cast<long>(po) * Wo + pw] : 0.f;
    float o[8];
    for (int i = 0; i < 8; ++i) o[i] = lrelu_fwd(fmaf(acc[i], d[static_cast<long>(n) * C + v * 8 + i], nz + bias[v * 8 + i]));
    store8(out + ((static_cast<long>(n) * Ho + po) * Wo + pw) * C + v * 8, o);
  }
}

template <typename T>
__global__ void blur_simple_bwd_kernel(const T* __restrict__ out, const T* __restrict__ gout, T* __restrict__ gT, const float* __restrict__ d,
                                       const float* __restrict__ noise, float noise_w, const float* __restrict__ bias, float* __restrict__ gdacc,
                                       const float* __restrict__ s_in, float* __restrict__ gs_in, int in_stride, int H, int W, int C) {
  const int n = blockIdx.y;
  const int vecs = C / 8, Ho = 2 * H, Wo = 2 * W, Hp = H + 1, Wp = W + 1, Hq = 2 * H + 2, Wq = 2 * W + 2;
  const long total = static_cast<long>(Hq) * Wq * vecs;
  const T* on = out + static_cast<long>(n) * Ho * Wo * C;
  const T* gn = gout + static_cast<long>(n) * Ho * Wo * C;
  T* gTn = gT + static_cast<long>(n) * 4 * Hp * Wp * C;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = idx % vecs;
    const long p = idx / vecs;
    const int r = p % Wq, q = p / Wq;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (q <= 2 * H && r <= 2 * W) {
      for (int t = 0; t < 4; ++t) {
        const int o = q - t + 1;
        if (o < 0 || o >= Ho) continue;
        for (int u = 0; u < 4; ++u) {
          const int pw = r - u + 1;
          if (pw < 0 || pw >= Wo) continue;
          const long off = (static_cast<long>(o) * Wo + pw) * C + v * 8;
          float ov[8], gv[8];
          load8(on + off, ov);
          load8(gn + off, gv);
          const float wgt = blur_w(t) * blur_w(u);
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt * lrelu_slope(ov[i]), gv[i], acc[i]);
        }
      }
      if (q < Ho && r < Wo) {
        const long off = (static_cast<long>(q) * Wo + r) * C + v * 8;
        float ov[8], gv[8];
        load8(on + off, ov);
        load8(gn + off, gv);
        const float nz = noise ? noise_w * noise[static_cast<long>(q) * Wo + r] : 0.f;
        for (int i = 0; i < 8; ++i) {
          const float si = s_in ? s_in[static_cast<long>(n) * in_stride + v * 8 + i] : 1.f;
          atomicAdd(gdacc + static_cast<long>(n) * C + v * 8 + i, si * gv[i] * lrelu_slope(ov[i]) * (lrelu_inv(ov[i]) - nz - bias[v * 8 + i]));
          if (gs_in) atomicAdd(gs_in + static_cast<long>(n) * in_stride + v * 8 + i, ov[i] * gv[i]);
        }
      }
    }
    for (int i = 0; i < 8; ++i) acc[i] *= d[static_cast<long>(n) * C + v * 8 + i] * (s_in ? s_in[static_cast<long>(n) * in_stride + v * 8 + i] : 1.f);
    store8(gTn + ((static_cast<long>((q & 1) * 2 + (r & 1)) * Hp + (q >> 1)) * Wp + (r >> 1)) * C + v * 8, acc);
  }
}

struct BlurGeom {
  int CV, bc, TC, pitch;   // channel vectors per position, column blocks per tile, tile columns, padded positions per row
};
__host__ __device__ inline int blur_cp(int r) { return r + (r >> 2); }
__host__ __device__ inline BlurGeom blur_geom(int vecs) {
  BlurGeom g;
  g.CV = vecs < 4 ? vecs : 4;
  g.bc = (256 / g.CV) / 8;
  g.TC = g.bc * 4;
  g.pitch = blur_cp(g.TC + 2) + 1;
  return g;
}

template <bool BWD>
__global__ void __launch_bounds__(256, 2) blur_tile_kernel(const bf16* __restrict__ src0 /* fwd: T (phase planar) ; bwd: out */,
                                                        const bf16* __restrict__ src1 /* bwd: gout */, bf16* __restrict__ dst,
                                                        const float* __restrict__ d, const float* __restrict__ noise, float noise_w,
                                                        const float* __restrict__ bias, float* __restrict__ gdacc, const float* __restrict__ s_in,
                                                        float* __restrict__ gs_in, int in_stride, int H, int W, int C) {
  extern __shared__ uint4 tile[];
  __shared__ float sred[32];
  const int n = blockIdx.y, cg = blockIdx.z;
  const int vecs = C / 8;
  const BlurGeom G = blur_geom(vecs);
  const int CV = G.CV, TC = G.TC, pitch = G.pitch, bc = G.bc;
  const int Ho = 2 * H, Wo = 2 * W, Hp = H + 1, Wp = W + 1;
  const int rows_out = BWD ? 2 * H + 2 : Ho;   // bwd writes every slot of every phase plane
  const int cols_out = BWD ? 2 * W + 2 : Wo;
  const int tiles_r = (rows_out + 15) / 16, tiles_c = (cols_out + TC - 1) / TC;
  const int v = threadIdx.x % CV, cvec = cg * CV + v;       // this thread's channel vector (fixed for the whole kernel)
  const int b = threadIdx.x / CV, jb = b % bc, ib = b / bc;  // its 2x4 block inside a tile
  float dv[8], bv[8], racc[8], si[8], rin[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    dv[c] = d[static_cast<long>(n) * C + cvec * 8 + c];
    bv[c] = bias[cvec * 8 + c];
    racc[c] = 0.f;
    rin[c] = 0.f;
    si[c] = (BWD && s_in) ? s_in[static_cast<long>(n) * in_stride + cvec * 8 + c] : 1.f;   // the gradient arrives without its style factor
  }
  const bf16* s0 = BWD ? src0 + static_cast<long>(n) * Ho * Wo * C : src0 + static_cast<long>(n) * 4 * Hp * Wp * C;
  const bf16* s1 = BWD ? src1 + static_cast<long>(n) * Ho * Wo * C : nullptr;
  bf16* dn = BWD ? dst + static_cast<long>(n) * 4 * Hp * Wp * C : dst + static_cast<long>(n) * Ho * Wo * C;
  const int halo = BWD ? 2 : 1;
  const int in_cols = TC + 3;
  const int n_in = 19 * in_cols * CV;
  for (int t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const int R0 = (t / tiles_c) * 16, C0 = (t % tiles_c) * TC;
    // ---- load phase
    for (int e = threadIdx.x; e < n_in; e += 256) {
      const int c = (e / CV) % in_cols, r = e / (CV * in_cols);
      const int q = R0 - halo + r, rr = C0 - halo + c;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (!BWD) {
        if (q >= 0 && q <= 2 * H && rr >= 0 && rr <= 2 * W) {
          const int plane = (q & 1) * 2 + (rr & 1);
          val = ldg8(s0 + ((static_cast<long>(plane) * Hp + (q >> 1)) * Wp + (rr >> 1)) * C + cvec * 8);
        }
      } else {
        if (q >= 0 && q < Ho && rr >= 0 && rr < Wo) {
          const long off = (static_cast<long>(q) * Wo + rr) * C + cvec * 8;
          float ov[8], gv[8];
          load8(s0 + off, ov);
          load8(s1 + off, gv);
          const bool mine = r >= 2 && r < 18 && c >= 2 && c < 2 + TC;   // the tile's own pixels: reductions
          if (mine) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) rin[cc] = fmaf(ov[cc], gv[cc], rin[cc]);
          }
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) gv[cc] *= si[cc] * lrelu_slope(ov[cc]);
          if (mine) {
            const float nz = noise ? noise_w * __ldg(noise + static_cast<long>(q) * Wo + rr) : 0.f;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) racc[cc] = fmaf(gv[cc], lrelu_inv(ov[cc]) - nz - bv[cc], racc[cc]);
          }
          val = pack8(gv);
        }
      }
      tile[(r * pitch + blur_cp(c)) * CV + v] = val;
    }
    __syncthreads();
    // ---- compute phase
    float acc[2][4][8];
    blur_block_2x4(2 * ib, 4 * jb, [&](int r, int c, float* f) { unpack8(tile[(r * pitch + blur_cp(c)) * CV + v], f); }, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int q = R0 + 2 * ib + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = C0 + 4 * jb + j;
        if (q >= rows_out || rr >= cols_out) continue;
        float o[8];
        if (!BWD) {
          const float nz = noise ? noise_w * __ldg(noise + static_cast<long>(q) * Wo + rr) : 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = lrelu_fwd(fmaf(acc[i][j][c], dv[c], nz + bv[c]));
          store8(dn + (static_cast<long>(q) * Wo + rr) * C + cvec * 8, o);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = acc[i][j][c] * dv[c];
          const int plane = (q & 1) * 2 + (rr & 1);
          store8(dn + ((static_cast<long>(plane) * Hp + (q >> 1)) * Wp + (rr >> 1)) * C + cvec * 8, o);
        }
      }
    }
    __syncthreads();
  }
  if (BWD) {
    // per-channel totals: threads with the same v (tid % CV) reduce through shared memory, then 8*CV atomics per block
    float* facc = reinterpret_cast<float*>(tile);   // reuse: [CV*8]
    if (threadIdx.x < CV * 8) facc[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float x = racc[c];
      // lanes l and l^CV, l^2CV, ... share v: butterfly over the bits above log2(CV)
      for (int o = 16; o >= CV; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) < CV) atomicAdd(&facc[v * 8 + c], x);
    }
    __syncthreads();
    if (threadIdx.x < CV * 8) atomicAdd(gdacc + static_cast<long>(n) * C + cg * CV * 8 + threadIdx.x, facc[threadIdx.x]);
    if (gs_in != nullptr) {   // style gradient of the conv that consumes this layer's output: sum x * gx~
      __syncthreads();
      if (threadIdx.x < CV * 8) facc[threadIdx.x] = 0.f;
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float x = rin[c];
        for (int o = 16; o >= CV; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) < CV) atomicAdd(&facc[v * 8 + c], x);
      }
      __syncthreads();
      if (threadIdx.x < CV * 8) atomicAdd(gs_in + static_cast<long>(n) * in_stride + cg * CV * 8 + threadIdx.x, facc[threadIdx.x]);
    }
  }
  (void)sred;
}

// The activation is piecewise linear through 0, so with m = act'(out) the pre-activation is out / m and
//   gy * y = (g m)(out/m - nz - b) = g*out - gy (nz + b):
// the first term is the same sum the consumer conv's style gradient needs (rin), the second is accumulated on the OUTPUT value
// gz = gy d and divided by d (> 0) once at the end.  Six ALU ops per element instead of eleven.
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ out, const T* gout, T* gz /* may alias gout (in place) */, const float* __restrict__ d,
                               const float* __restrict__ noise, float noise_w, const float* __restrict__ bias, float* __restrict__ gdacc,
                               const float* __restrict__ s_in, float* __restrict__ gs_in, int vec_stride, int HW, int C) {
  extern __shared__ float sacc[];
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const long total = static_cast<long>(HW) * vecs;
  const long base = static_cast<long>(n) * HW * C;
  const int cv = threadIdx.x % vecs;
  float dv[8], bv[8], racc[8], si[8], rin[8], kp[8], kn[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dv[i] = d[static_cast<long>(n) * C + cv * 8 + i];
    bv[i] = bias[cv * 8 + i];
    racc[i] = 0.f;
    si[i] = s_in ? s_in[static_cast<long>(n) * vec_stride + cv * 8 + i] : 1.f;
    rin[i] = 0.f;
    kp[i] = si[i] * dv[i] * SFK_SQRT2;          // gz = g * (out > 0 ? kp : kn)
    kn[i] = kp[i] * 0.2f;
  }
  // two pixels per iteration: both loads are in flight before the first is consumed (the kernel is bound by memory latency x
  // occupancy, not by bandwidth, at one 32-byte request per thread)
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += 2 * stride) {
    const long pA = idx / vecs, pB = (idx + stride) / vecs;
    const bool hasB = idx + stride < total;
    const long offA = base + pA * C + cv * 8, offB = base + pB * C + cv * 8;
    float ovA[8], gvA[8], ovB[8], gvB[8], o[8];
    load8(out + offA, ovA);
    load8p(gout + offA, gvA);   // gz may alias gout (in-place)
    float nzA = noise ? __ldg(noise + pA) : 0.f, nzB = 0.f;
    if (hasB) {
      load8(out + offB, ovB);
      load8p(gout + offB, gvB);
      nzB = noise ? __ldg(noise + pB) : 0.f;
    }
    nzA *= noise_w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      rin[i] = fmaf(ovA[i], gvA[i], rin[i]);
      o[i] = gvA[i] * (ovA[i] > 0.f ? kp[i] : kn[i]);
      racc[i] = fmaf(o[i], nzA + bv[i], racc[i]);
    }
    store8(gz + offA, o);
    if (hasB) {
      nzB *= noise_w;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rin[i] = fmaf(ovB[i], gvB[i], rin[i]);
        o[i] = gvB[i] * (ovB[i] > 0.f ? kp[i] : kn[i]);
        racc[i] = fmaf(o[i], nzB + bv[i], racc[i]);
      }
      store8(gz + offB, o);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) racc[i] = si[i] * rin[i] - racc[i] / dv[i];
  flush_channel_acc(sacc, racc, cv, C, gdacc + static_cast<long>(n) * C);
  if (gs_in != nullptr) {
    __syncthreads();
    flush_channel_acc(sacc, rin, cv, C, gs_in + static_cast<long>(n) * vec_stride);
  }
}

// ============================================================================================
// ToRGB (1x1 modulated conv to 3 channels, no demod) + skip upsample
__device__ __forceinline__ float skip_up_sample(const float* __restrict__ sk, int hs, int ws, int o, int p) {
  // upfirdn2d(skip, k*4, up=2, pad=(2,1)): even o -> (o/2-1: 1/4, o/2: 3/4); odd o -> ((o-1)/2: 3/4, (o+1)/2: 1/4)
  int i0, i1, j0, j1;
  float a0, a1, b0, b1;
  if ((o & 1) == 0) { i0 = o / 2 - 1; i1 = o / 2; a0 = 0.25f; a1 = 0.75f; } else { i0 = (o - 1) / 2; i1 = (o + 1) / 2; a0 = 0.75f; a1 = 0.25f; }
  if ((p & 1) == 0) { j0 = p / 2 - 1; j1 = p / 2; b0 = 0.25f; b1 = 0.75f; } else { j0 = (p - 1) / 2; j1 = (p + 1) / 2; b0 = 0.75f; b1 = 0.25f; }
  float r = 0.f;
  const bool vi0 = i0 >= 0 && i0 < hs, vi1 = i1 >= 0 && i1 < hs, vj0 = j0 >= 0 && j0 < ws, vj1 = j1 >= 0 && j1 < ws;
  if (vi0 && vj0) r += a0 * b0 * __ldg(sk + static_cast<long>(i0) * ws + j0);
  if (vi0 && vj1) r += a0 * b1 * __ldg(sk + static_cast<long>(i0) * ws + j1);
  if (vi1 && vj0) r += a1 * b0 * __ldg(sk + static_cast<long>(i1) * ws + j0);
  if (vi1 && vj1) r += a1 * b1 * __ldg(sk + static_cast<long>(i1) * ws + j1);
  return r;
}

// LP lanes cooperate on one pixel (LP = max(1, C/128)), each lane owns channel vectors lane, lane+LP, ...; per-pixel work (index
// math, skip upsample, planar stores) is done once, so few lanes per pixel is right for the narrow high-resolution layers.
// Modulated weights sit in shared memory as [vector][colour][8]: six broadcast LDS.128 feed the 24 FMAs of one 16-byte load.
template <typename T>
__global__ void torgb_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wrgb, const float* __restrict__ s, int s_stride,
                                 const float* __restrict__ bias, const float* __restrict__ skip, float* __restrict__ rgb, int H, int W, int C, int LP) {
  extern __shared__ __align__(16) float swm[];  // [C/8][3][8] modulated weights of this sample
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    const int col = i / C, c = i % C;
    swm[((c >> 3) * 3 + col) * 8 + (c & 7)] = wrgb[i] * s[static_cast<long>(n) * s_stride + c];
  }
  __syncthreads();
  const int vecs = C / 8;
  const long HW = static_cast<long>(H) * W;
  const long total = HW * LP;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  const long rounds = (total + stride - 1) / stride;
  for (long rr = 0; rr < rounds; ++rr) {
    const long idx = rr * stride + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const bool active = idx < total;
    const int sub = idx % LP;
    const long p = idx / LP;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (active) {
      const T* xp = x + (static_cast<long>(n) * HW + p) * C;
      auto dot = [&](const float (&xv)[8], int v) {
        const float4* wp = reinterpret_cast<const float4*>(swm + v * 24);
        const float4 r0 = wp[0], r1 = wp[1], g0 = wp[2], g1 = wp[3], b0 = wp[4], b1 = wp[5];
        a0 = fmaf(xv[0], r0.x, a0); a0 = fmaf(xv[1], r0.y, a0); a0 = fmaf(xv[2], r0.z, a0); a0 = fmaf(xv[3], r0.w, a0);
        a0 = fmaf(xv[4], r1.x, a0); a0 = fmaf(xv[5], r1.y, a0); a0 = fmaf(xv[6], r1.z, a0); a0 = fmaf(xv[7], r1.w, a0);
        a1 = fmaf(xv[0], g0.x, a1); a1 = fmaf(xv[1], g0.y, a1); a1 = fmaf(xv[2], g0.z, a1); a1 = fmaf(xv[3], g0.w, a1);
        a1 = fmaf(xv[4], g1.x, a1); a1 = fmaf(xv[5], g1.y, a1); a1 = fmaf(xv[6], g1.z, a1); a1 = fmaf(xv[7], g1.w, a1);
        a2 = fmaf(xv[0], b0.x, a2); a2 = fmaf(xv[1], b0.y, a2); a2 = fmaf(xv[2], b0.z, a2); a2 = fmaf(xv[3], b0.w, a2);
        a2 = fmaf(xv[4], b1.x, a2); a2 = fmaf(xv[5], b1.y, a2); a2 = fmaf(xv[6], b1.z, a2); a2 = fmaf(xv[7], b1.w, a2);
      };
      int v = sub;
      for (; v + 3 * LP < vecs; v += 4 * LP) {   // four 16-byte loads in flight per thread
        float x0[8], x1[8], x2[8], x3[8];
        load8(xp + v * 8, x0);
        load8(xp + (v + LP) * 8, x1);
        load8(xp + (v + 2 * LP) * 8, x2);
        load8(xp + (v + 3 * LP) * 8, x3);
        dot(x0, v);
        dot(x1, v + LP);
        dot(x2, v + 2 * LP);
        dot(x3, v + 3 * LP);
      }
      for (; v < vecs; v += LP) {
        float xv[8];
        load8(xp + v * 8, xv);
        dot(xv, v);
      }
    }
    for (int o = LP >> 1; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (active && sub == 0) {
      const int h = p / W, w = p % W;
      float r0 = a0 + bias[0], r1 = a1 + bias[1], r2 = a2 + bias[2];
      if (skip) {
        const int hs = H / 2, ws = W / 2;
        const float* sk = skip + static_cast<long>(n) * 3 * hs * ws;
        r0 += skip_up_sample(sk, hs, ws, h, w);
        r1 += skip_up_sample(sk + static_cast<long>(hs) * ws, hs, ws, h, w);
        r2 += skip_up_sample(sk + 2L * hs * ws, hs, ws, h, w);
      }
      float* o = rgb + static_cast<long>(n) * 3 * HW + p;
      o[0] = r0;
      o[HW] = r1;
      o[2 * HW] = r2;
    }
  }
}

template <typename T>
__global__ void torgb_bwd_kernel(const T* __restrict__ x, const float* __restrict__ wrgb, const float* __restrict__ s, int s_stride,
                                 const float* __restrict__ grgb, T* __restrict__ gx, float* __restrict__ gs, int gs_stride, int HW, int C) {
  extern __shared__ float sm[];  // [3][C] weights, then [C] accumulators
  float* sw = sm;
  float* sacc = sm + 3 * C;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sw[i] = wrgb[i];
  __syncthreads();
  const int vecs = C / 8;
  const int cv = threadIdx.x % vecs;
  float sv[8], racc[8], w0[8], w1[8], w2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sv[i] = s[static_cast<long>(n) * s_stride + cv * 8 + i];
    racc[i] = 0.f;
    w0[i] = sw[cv * 8 + i];
    w1[i] = sw[C + cv * 8 + i];
    w2[i] = sw[2 * C + cv * 8 + i];
  }
  const long total = static_cast<long>(HW) * vecs;
  const float* g0 = grgb + static_cast<long>(n) * 3 * HW;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long p = idx / vecs;
    const float ga = __ldg(g0 + p), gb = __ldg(g0 + HW + p), gc = __ldg(g0 + 2L * HW + p);
    const long off = (static_cast<long>(n) * HW + p) * C + cv * 8;
    float xv[8], o[8];
    load8(x + off, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gt = w0[i] * ga + w1[i] * gb + w2[i] * gc;
      racc[i] = fmaf(xv[i], gt, racc[i]);
      o[i] = sv[i] * gt;
    }
    store8(gx + off, o);
  }
  flush_channel_acc(sacc, racc, cv, C, gs + static_cast<long>(n) * gs_stride);
}

// ToRGB backward + activation backward of the conv that feeds it, in one pass over the conv's output / gradient buffers:
//   g   = gin (gradient already written by the next resolution's upsample conv; absent at the top) + s_rgb * (wrgb^T grgb)
//   gs_rgb[n][i] += sum_hw out * (wrgb^T grgb)                      (ToRGB style gradient)
//   gy  = g * act'(out);  gdacc[n][j] += sum_hw gy * y;  gz = gy * d      (as act_bwd)
template <typename T>
__global__ void __launch_bounds__(kBlock, 2)
act_torgb_bwd_kernel(const T* __restrict__ out, const T* gin, T* gz, const float* __restrict__ d, const float* __restrict__ noise, float noise_w,
                     const float* __restrict__ bias, float* __restrict__ gdacc, const float* __restrict__ wrgb, const float* __restrict__ s,
                     int s_stride, const float* __restrict__ grgb, float* __restrict__ gs, int gs_stride, const float* __restrict__ s_in,
                     float* __restrict__ gs_in, int HW, int C) {
  extern __shared__ __align__(16) float sm[];   // [C] accumulator scratch, then ToRGB weights as [C/8][3][8]
  float* sacc = sm;
  float* sw = sm + C;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    const int col = i / C, c = i % C;
    sw[((c >> 3) * 3 + col) * 8 + (c & 7)] = wrgb[i];
  }
  __syncthreads();
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int cv = threadIdx.x % vecs;
  const float4* wq = reinterpret_cast<const float4*>(sw + cv * 24);
  float sv[8], bv[8], racc[8], rrgb[8], si[8], rin[8], kp[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cv * 8 + i;
    sv[i] = s[static_cast<long>(n) * s_stride + c];
    bv[i] = bias[c];
    racc[i] = 0.f;
    rrgb[i] = 0.f;
    si[i] = s_in ? s_in[static_cast<long>(n) * s_stride + c] : 1.f;
    rin[i] = 0.f;
    kp[i] = d[static_cast<long>(n) * C + c] * SFK_SQRT2;      // gz = g * kp * (out > 0 ? 1 : 0.2)
  }
  const long total = static_cast<long>(HW) * vecs;
  const long base = static_cast<long>(n) * HW * C;
  const float* g0 = grgb + static_cast<long>(n) * 3 * HW;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  auto body = [&](const float (&ov)[8], const float (&gv)[8], float ga, float gb, float gc, float nz, long off) {
    const float4 r0 = wq[0], r1 = wq[1], q0 = wq[2], q1 = wq[3], b0 = wq[4], b1 = wq[5];
    const float w0[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const float w1[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const float w2[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gt = w0[i] * ga + w1[i] * gb + w2[i] * gc;
      rrgb[i] = fmaf(ov[i], gt, rrgb[i]);
      rin[i] = fmaf(ov[i], gv[i], rin[i]);
      const float g = fmaf(sv[i], gt, gv[i] * si[i]);
      o[i] = g * kp[i] * (ov[i] > 0.f ? 1.f : 0.2f);
      racc[i] = fmaf(o[i], nz + bv[i], racc[i]);
    }
    store8(gz + off, o);
  };
  // two pixels per iteration so that both sets of loads are in flight together (see act_bwd_kernel)
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += 2 * stride) {
    const long pA = idx / vecs, pB = (idx + stride) / vecs;
    const bool hasB = idx + stride < total;
    const long offA = base + pA * C + cv * 8, offB = base + pB * C + cv * 8;
    float ovA[8], gvA[8], ovB[8], gvB[8];
    load8(out + offA, ovA);
    if (gin != nullptr) {
      load8p(gin + offA, gvA);   // gz may alias gin (in-place)
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) gvA[i] = 0.f;
    }
    const float gaA = __ldg(g0 + pA), gbA = __ldg(g0 + HW + pA), gcA = __ldg(g0 + 2L * HW + pA);
    const float nzA = noise ? __ldg(noise + pA) : 0.f;
    float gaB = 0.f, gbB = 0.f, gcB = 0.f, nzB = 0.f;
    if (hasB) {
      load8(out + offB, ovB);
      if (gin != nullptr) {
        load8p(gin + offB, gvB);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) gvB[i] = 0.f;
      }
      gaB = __ldg(g0 + pB); gbB = __ldg(g0 + HW + pB); gcB = __ldg(g0 + 2L * HW + pB);
      nzB = noise ? __ldg(noise + pB) : 0.f;
    }
    body(ovA, gvA, gaA, gbA, gcA, noise_w * nzA, offA);
    if (hasB) body(ovB, gvB, gaB, gbB, gcB, noise_w * nzB, offB);
  }
  // gy*y = g*out - gy*(nz+b)  (see act_bwd_kernel);  sum g*out = s_in*rin + s_rgb*rrgb;  racc was accumulated on gz = gy*d
#pragma unroll
  for (int i = 0; i < 8; ++i) racc[i] = si[i] * rin[i] + sv[i] * rrgb[i] - racc[i] * SFK_SQRT2 / kp[i];
  flush_channel_acc(sacc, racc, cv, C, gdacc + static_cast<long>(n) * C);
  __syncthreads();
  flush_channel_acc(sacc, rrgb, cv, C, gs + static_cast<long>(n) * gs_stride);
  if (gs_in != nullptr) {
    __syncthreads();
    flush_channel_acc(sacc, rin, cv, C, gs_in + static_cast<long>(n) * gs_stride);
  }
}

// two horizontally adjacent outputs per thread: per input row one aligned float4 and two scalars instead of eight scalar loads,
// separable 4-tap filter (needs W % 4 == 0; the scalar kernel below covers everything else)
__global__ void rgb_down2_kernel(const float* __restrict__ g, float* __restrict__ gs, long planes, int H, int W) {
  const int hs = H / 2, ws = W / 2, wp = ws / 2;
  const long total = planes * hs * wp;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int jp = idx % wp;
    long p = idx / wp;
    const int i = p % hs;
    const long pl = p / hs;
    const float* gp = g + pl * H * W;
    const int c0 = 4 * jp;                 // input columns c0-1 .. c0+4 feed outputs 2jp and 2jp+1
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int o = 2 * i + t - 1;
      if (o < 0 || o >= H) continue;
      const float* row = gp + static_cast<long>(o) * W;
      const float4 m = __ldg(reinterpret_cast<const float4*>(row + c0));
      const float l = c0 > 0 ? __ldg(row + c0 - 1) : 0.f;
      const float r = c0 + 4 < W ? __ldg(row + c0 + 4) : 0.f;
      const float h0 = blur_w(0) * l + blur_w(1) * m.x + blur_w(2) * m.y + blur_w(3) * m.z;
      const float h1 = blur_w(0) * m.y + blur_w(1) * m.z + blur_w(2) * m.w + blur_w(3) * r;
      a0 = fmaf(blur_w(t), h0, a0);
      a1 = fmaf(blur_w(t), h1, a1);
    }
    *reinterpret_cast<float2*>(gs + (pl * hs + i) * ws + 2 * jp) = make_float2(a0, a1);
  }
}

__global__ void rgb_down_kernel(const float* __restrict__ g, float* __restrict__ gs, long planes, int H, int W) {
  // gskip[i][j] = sum_{t,u} g[2i+t-1][2j+u-1] k'[t] k'[u]
  const int hs = H / 2, ws = W / 2;
  const long total = planes * hs * ws;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int j = idx % ws;
    long p = idx / ws;
    const int i = p % hs;
    const long pl = p / hs;
    const float* gp = g + pl * H * W;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int o = 2 * i + t - 1;
      if (o < 0 || o >= H) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = 2 * j + u - 1;
        if (q < 0 || q >= W) continue;
        acc = fmaf(blur_w(t) * blur_w(u), __ldg(gp + static_cast<long>(o) * W + q), acc);
      }
    }
    gs[idx] = acc;
  }
}

// ============================================================================================
// small dense helpers
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias, float* __restrict__ y,
                                  int N, int In, int Out) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < Out; o += warps) {
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = lane; k < In; k += 32) acc = fmaf(__ldg(Wt + static_cast<long>(o) * In + k), __ldg(x + static_cast<long>(n) * In + k), acc);
      acc = warp_sum(acc);
      if (lane == 0) y[static_cast<long>(n) * Out + o] = acc + (bias ? bias[o] : 0.f);
    }
  }
}

// split over the output dimension: block (x = chunk of 128 inputs, y = slice of 32 outputs) reads its W rows ONCE and serves
// every sample from registers (the slice of gy sits in shared memory); partial sums are merged with atomics (gx is zeroed by the
// launcher).  Many small slices: the kernel is a single pass over W and needs the whole GPU's load parallelism.
constexpr int kLinSlice = 32;
__global__ void linear_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ Wt, float* __restrict__ gx, int N, int In, int Out) {
  __shared__ float sg[16][kLinSlice];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int o0 = blockIdx.y * kLinSlice;
  const int no = min(kLinSlice, Out - o0);
  for (int nb = 0; nb < N; nb += 16) {
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * kLinSlice; t += blockDim.x) {
      const int q = t / kLinSlice, o = t % kLinSlice;
      sg[q][o] = (nb + q < N && o < no) ? __ldg(gy + static_cast<long>(nb + q) * Out + o0 + o) : 0.f;
    }
    __syncthreads();
    if (i >= In) continue;
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
#pragma unroll 4
    for (int o = 0; o < no; ++o) {
      const float wv = __ldg(Wt + static_cast<long>(o0 + o) * In + i);
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = fmaf(sg[q][o], wv, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (nb + q < N) atomicAdd(gx + static_cast<long>(nb + q) * In + i, acc[q]);
  }
}

__global__ void zero_kernel(float* p, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) p[i] = 0.f;
}

__global__ void fuse_spatial_fwd_kernel(const float* sa, const float* sb, const float* al, const float* be, const float* c, float* s, int N, int D) {
  const long total = static_cast<long>(N) * D;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = idx % D;
    const float a = sa[idx], b = sb[idx];
    const float q = 1.f / (1.f + __expf(-(al[r] * a + be[r] * b + c[r])));
    s[idx] = q * a + (1.f - q) * b;
  }
}

__global__ void fuse_spatial_bwd_kernel(const float* sa, const float* sb, const float* al, const float* be, const float* c, const float* gs,
                                        float* gsa, float* gsb, int N, int D) {
  const long total = static_cast<long>(N) * D;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = idx % D;
    const float a = sa[idx], b = sb[idx], g = gs[idx];
    const float q = 1.f / (1.f + __expf(-(al[r] * a + be[r] * b + c[r])));
    const float gq = g * (a - b) * q * (1.f - q);
    gsa[idx] = g * q + gq * al[r];
    gsb[idx] = g * (1.f - q) + gq * be[r];
  }
}

__global__ void axpby_kernel(const float* x, const float* y, float* out, float a, float b, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int N, int C, int H, int W) {
  const long total = static_cast<long>(N) * H * W * C;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = idx % C;
    long p = idx / C;
    const int w = p % W;
    p /= W;
    const int h = p % H;
    const int n = p / H;
    from_f32(&y[idx], x[((static_cast<long>(n) * C + c) * H + h) * W + w]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W) {
  const long total = static_cast<long>(N) * H * W * C;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int w = idx % W;
    long p = idx / W;
    const int h = p % H;
    p /= H;
    const int c = p % C;
    const int n = p / C;
    y[idx] = to_f32(x[((static_cast<long>(n) * H + h) * W + w) * C + c]);
  }
}

// ============================================================================================
// perturbation updates
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

__device__ __forceinline__ void block_add(float v, float* dst) {
  v = warp_sum(v);
  __shared__ float red[kBlock / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0 && dst) {
    float s = 0.f;
    for (int q = 0; q < blockDim.x / 32; ++q) s += red[q];
    atomicAdd(dst, s);
  }
  __syncthreads();
}

__device__ __forceinline__ float pooled_grad(const float* __restrict__ gpool, int n, long i, int S, int k) {
  const int Sp = S / k;
  const int w = i % S;
  long p = i / S;
  const int h = p % S;
  const int c = p / S;
  return __ldg(gpool + ((static_cast<long>(n) * 3 + c) * Sp + h / k) * Sp + w / k);
}

// 4 consecutive pixels of one row per thread (float4): S % 4 == 0 and (k == 1 or 4 % k == 0 or k % 4 == 0) hold for S = 2^m
__global__ void update_linf_kernel(float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ gpool, float alpha, float eps,
                                   float dir, float lo, float hi, float* __restrict__ stats, int S, int k) {
  const int n = blockIdx.y;
  const long per4 = 3L * S * S / 4;
  const int Sp = S / k, S4 = S / 4;
  float dsum = 0.f;
  float4* xv = reinterpret_cast<float4*>(x + static_cast<long>(n) * 3 * S * S);
  const float4* x0v = reinterpret_cast<const float4*>(x0 + static_cast<long>(n) * 3 * S * S);
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int w4 = i % S4;
    long p = i / S4;
    const int h = p % S;
    const int c = p / S;
    const float* gr = gpool + ((static_cast<long>(n) * 3 + c) * Sp + h / k) * Sp;
    const float4 c0 = __ldg(x0v + i);
    float4 v = xv[i];
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = __ldg(gr + (w4 * 4 + j) / k);
    float* vp = reinterpret_cast<float*>(&v);
    const float* cp = reinterpret_cast<const float*>(&c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = vp[j] + dir * alpha * sgn(g[j]);
      const float dl = fminf(fmaxf(t - cp[j], -eps), eps);
      t = fminf(fmaxf(cp[j] + dl, lo), hi);
      vp[j] = t;
      dsum += fabsf(t - cp[j]);
    }
    xv[i] = v;
  }
  block_add(dsum, stats ? stats + n : nullptr);
}

// PGD random start on the device (code/attack/interpolation.py:74-76): x = clamp(x0 + eps * U(-1,1), lo, hi) with a counter-based
// generator (one 64-bit SplitMix finaliser per float4: a pure function of (seed, element index), so the result does not depend on
// the grid, the rank layout or the replay count)
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void random_start_kernel(float4* __restrict__ x, const float4* __restrict__ x0, float eps, float lo, float hi, uint64_t seed, long n4) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint64_t a = splitmix64(seed ^ (static_cast<uint64_t>(i) * 2ull)), b = splitmix64(seed ^ (static_cast<uint64_t>(i) * 2ull + 1ull));
    const float sc = eps * (2.0f / 16777216.0f);       // 24-bit mantissa uniforms in [0,1) -> [-eps, eps)
    const float4 c = __ldg(x0 + i);
    float4 v;
    v.x = fminf(fmaxf(c.x + (static_cast<float>(static_cast<uint32_t>(a) >> 8) * sc - eps), lo), hi);
    v.y = fminf(fmaxf(c.y + (static_cast<float>(static_cast<uint32_t>(a >> 32) >> 8) * sc - eps), lo), hi);
    v.z = fminf(fmaxf(c.z + (static_cast<float>(static_cast<uint32_t>(b) >> 8) * sc - eps), lo), hi);
    v.w = fminf(fmaxf(c.w + (static_cast<float>(static_cast<uint32_t>(b >> 32) >> 8) * sc - eps), lo), hi);
    x[i] = v;
  }
}

__global__ void update_patch_kernel(float* __restrict__ x, const float* __restrict__ x0, float* __restrict__ patch, const float* __restrict__ mask,
                                    const float* __restrict__ gpool, float lr, float dir, int use_sign, const float* __restrict__ lo,
                                    const float* __restrict__ hi, float gscale, float* __restrict__ stats, int S, int k) {
  const int n = blockIdx.y;
  const long per = 3L * S * S;
  float dsum = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long off = static_cast<long>(n) * per + i;
    const float g = gscale * pooled_grad(gpool, n, i, S, k);
    const float pv = patch[off] + dir * lr * (use_sign ? sgn(g) : g);
    patch[off] = pv;
    const float m = mask[off], c0 = x0[off];
    float v = (1.f - m) * c0 + m * pv;
    v = fminf(fmaxf(v, lo[n]), hi[n]);
    x[off] = v;
    dsum += fabsf(v - c0);
  }
  block_add(dsum, stats ? stats + n : nullptr);
}

__global__ void update_adam_kernel(float* __restrict__ x, const float* __restrict__ gpool, const float* __restrict__ gfull, float gfull_scale,
                                   float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2, float eps, float bc1, float bc2,
                                   float gscale, int S, int k) {
  const int n = blockIdx.y;
  const long per = 3L * S * S;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long off = static_cast<long>(n) * per + i;
    float g = gscale * pooled_grad(gpool, n, i, S, k);
    if (gfull) g += gfull_scale * gfull[off];
    const float mm = b1 * m[off] + (1.f - b1) * g;
    const float vv = b2 * v[off] + (1.f - b2) * g * g;
    m[off] = mm;
    v[off] = vv;
    x[off] -= lr * (mm / bc1) / (sqrtf(vv / bc2) + eps);
  }
}

__global__ void update_l2_kernel(float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ gpool, float* __restrict__ norms,
                                 float* __restrict__ dn, float alpha, float eps, float dir, float lo, float hi, int phase, int S, int k) {
  const int n = blockIdx.y;
  const long per = 3L * S * S;
  float acc = 0.f;
  const float gn = phase == 1 ? fmaxf(sqrtf(norms[n]), 1e-12f) : 1.f;
  const float sc = phase == 2 ? fminf(eps / fmaxf(sqrtf(dn[n]), 1e-12f), 1.f) : 1.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long off = static_cast<long>(n) * per + i;
    if (phase == 0) {
      const float g = pooled_grad(gpool, n, i, S, k);
      acc += g * g;
    } else if (phase == 1) {
      const float g = pooled_grad(gpool, n, i, S, k);
      const float v = x[off] + dir * alpha * g / gn;
      x[off] = v;
      const float dl = v - x0[off];
      acc += dl * dl;
    } else {
      const float c0 = x0[off];
      x[off] = fminf(fmaxf(c0 + (x[off] - c0) * sc, lo), hi);
    }
  }
  if (phase == 0) block_add(acc, norms + n);
  if (phase == 1) block_add(acc, dn + n);
}

__global__ void minmax_kernel(const float* __restrict__ x, float* __restrict__ lo, float* __restrict__ hi, long per) {
  const int n = blockIdx.x;
  float mn = INFINITY, mx = -INFINITY;
  for (long i = threadIdx.x; i < per; i += blockDim.x) {
    const float v = x[static_cast<long>(n) * per + i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[32], smx[32];
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < blockDim.x / 32; ++q) {
      mn = fminf(mn, smn[q]);
      mx = fmaxf(mx, smx[q]);
    }
    lo[n] = mn;
    hi[n] = mx;
  }
}

// ---- universal (shared) patch: one patch for every image of the batch -------------------------------------
// gsum[c][h][w] = mask * gscale * sum_n gpool[n][c][h/k][w/k]: the batch's gradient w.r.t. the shared patch (each image sees the
// patch through its own mask apply, adversarial_patch.py:137).  Overwrites gsum; one thread per patch pixel, images in registers.
__global__ void patch_grad_reduce_kernel(const float* __restrict__ gpool, const float* __restrict__ mask, float* __restrict__ gsum,
                                         float gscale, int N, int S, int k) {
  const long per = 3L * S * S;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float m = __ldg(mask + i);
    float acc = 0.f;
    if (m != 0.f) {
      for (int n = 0; n < N; ++n) acc += pooled_grad(gpool, n, i, S, k);
    }
    gsum[i] = m * gscale * acc;
  }
}

// x[n] = clamp((1 - mask) * x0[n] + mask * patch, lo[n], hi[n]) with ONE patch / mask for all images (attack_main2.py:416-418)
__global__ void patch_apply_shared_kernel(float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ patch,
                                          const float* __restrict__ mask, const float* __restrict__ lo, const float* __restrict__ hi, int S) {
  const int n = blockIdx.y;
  const long per = 3L * S * S;
  const float l = lo[n], h = hi[n];
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < per; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long off = static_cast<long>(n) * per + i;
    const float m = __ldg(mask + i);
    const float v = (1.f - m) * x0[off] + m * __ldg(patch + i);
    x[off] = fminf(fmaxf(v, l), h);
  }
}

// ---- SSIM (cal_SSMI, interpolation.py:903-919: skimage.metrics.structural_similarity of the rgb2gray images, library defaults:
// 7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, mean over the pixels whose window lies inside the image).
// One thread per window centre; the two gray tiles (16+6)^2 are built in shared memory from the NCHW fp32 images.
constexpr int kSsimT = 16, kSsimW = 7, kSsimR = 3;
__global__ void __launch_bounds__(kSsimT * kSsimT) ssim_gray7_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                                    int H, int W, float c1, float c2, float inv_count) {
  __shared__ float ta[kSsimT + 2 * kSsimR][kSsimT + 2 * kSsimR + 1];
  __shared__ float tb[kSsimT + 2 * kSsimR][kSsimT + 2 * kSsimR + 1];
  const int n = blockIdx.z;
  const int h0 = blockIdx.y * kSsimT, w0 = blockIdx.x * kSsimT;     // window centres (h0 + ty + R, w0 + tx + R)
  const long plane = static_cast<long>(H) * W;
  const float* pa = a + static_cast<long>(n) * 3 * plane;
  const float* pb = b + static_cast<long>(n) * 3 * plane;
  constexpr int TT = kSsimT + 2 * kSsimR;
  for (int i = threadIdx.x; i < TT * TT; i += blockDim.x) {
    const int r = i / TT, c = i % TT;
    const int h = h0 + r, w = w0 + c;
    float ga = 0.f, gb = 0.f;
    if (h < H && w < W) {
      const long o = static_cast<long>(h) * W + w;
      ga = 0.2125f * __ldg(pa + o) + 0.7154f * __ldg(pa + plane + o) + 0.0721f * __ldg(pa + 2 * plane + o);   // skimage.color.rgb2gray
      gb = 0.2125f * __ldg(pb + o) + 0.7154f * __ldg(pb + plane + o) + 0.0721f * __ldg(pb + 2 * plane + o);
    }
    ta[r][c] = ga;
    tb[r][c] = gb;
  }
  __syncthreads();
  const int ty = threadIdx.x / kSsimT, tx = threadIdx.x % kSsimT;
  float val = 0.f;
  if (h0 + ty + 2 * kSsimR < H && w0 + tx + 2 * kSsimR < W) {     // whole window inside the image
    // second moments about the window's centre pixel: (co)variances are shift invariant, and E[x^2] - E[x]^2 of the raw values would
    // cancel catastrophically in fp32 on the smooth regions where SSIM is most sensitive
    const float u0 = ta[ty + kSsimR][tx + kSsimR], v0 = tb[ty + kSsimR][tx + kSsimR];
    float sa = 0.f, sb = 0.f, saa = 0.f, sbb = 0.f, sab = 0.f;
#pragma unroll
    for (int r = 0; r < kSsimW; ++r)
#pragma unroll
      for (int c = 0; c < kSsimW; ++c) {
        const float u = ta[ty + r][tx + c] - u0, v = tb[ty + r][tx + c] - v0;
        sa += u; sb += v; saa = fmaf(u, u, saa); sbb = fmaf(v, v, sbb); sab = fmaf(u, v, sab);
      }
    constexpr float NP = kSsimW * kSsimW, cov_norm = NP / (NP - 1.f);
    const float du = sa / NP, dv = sb / NP;
    const float ux = u0 + du, uy = v0 + dv;
    const float vx = cov_norm * (saa / NP - du * du), vy = cov_norm * (sbb / NP - dv * dv), vxy = cov_norm * (sab / NP - du * dv);
    val = ((2.f * ux * uy + c1) * (2.f * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
  }
  val = warp_sum(val);
  __shared__ float red[kSsimT * kSsimT / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int q = 0; q < kSsimT * kSsimT / 32; ++q) s += red[q];
    atomicAdd(out + n, s * inv_count);
  }
}

inline cudaStream_t S_(sfk_stream_t s) { return static_cast<cudaStream_t>(s); }
inline unsigned per_sample_blocks(long items, int n) {
  long b = (items + kBlock - 1) / kBlock;
  long cap = (static_cast<long>(sfk_num_sms()) * 8 + n - 1) / n;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

inline unsigned per_sample_blocks128(long items, int n) {
  long b = (items + 127) / 128;
  long cap = (static_cast<long>(sfk_num_sms()) * 12 + n - 1) / n;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace

// =============================================================================================
extern "C" {

int sfk_conv_c3_fwd(const float* x, const float* w, const float* bias, void* out, int n, int h, int w_, int cout, int relu, sfk_stream_t s) {
  SFK_REQUIRE(x && w && out, SFK_E_ARG, "conv_c3_fwd: null");
  SFK_REQUIRE(cout % 8 == 0 && cout <= 512, SFK_E_SHAPE, "conv_c3_fwd: cout must be a multiple of 8, <= 512");
  const size_t smem = static_cast<size_t>(28 * cout) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(conv_c3_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(conv_c3_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      conv_c3_fwd_kernel<T><<<grid_for(static_cast<long>(n) * h * ((w_ + 3) / 4) * (cout / 8)), kBlock, smem, S_(s)>>>(x, w, bias, static_cast<T*>(out), n, h, w_, cout, relu);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("conv_c3_fwd");
}

int sfk_conv_c3_bwd(const void* g, const float* w, float* gx, int n, int h, int w_, int cout, sfk_stream_t s) {
  SFK_REQUIRE(g && w && gx, SFK_E_ARG, "conv_c3_bwd: null");
  SFK_REQUIRE(cout % 8 == 0 && cout <= 512, SFK_E_SHAPE, "conv_c3_bwd: cout must be a multiple of 8, <= 512");
  int lp = 1;
  while (lp < 8 && lp * 2 <= cout / 8) lp *= 2;
  const size_t smem = static_cast<size_t>(9 * (cout / 8) * 36) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(conv_c3_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(conv_c3_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      conv_c3_bwd_kernel<T><<<grid_for(static_cast<long>(n) * h * ((w_ + 3) / 4) * lp), kBlock, smem, S_(s)>>>(static_cast<const T*>(g), w, gx, n, h, w_, cout, lp);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("conv_c3_bwd");
}

int sfk_c3_pack(const float* x, void* xp, int n, int h, int w, sfk_stream_t s) {
  SFK_REQUIRE(x && xp && sfk_aligned16(xp) && n >= 1 && h >= 1 && w >= 1, SFK_E_ARG, "c3_pack: bad args");
  SFK_REQUIRE(!sfk_act_f32(), SFK_E_ARG, "c3_pack: bf16 storage only (the fp32 parity mode runs sfk_conv_c3_fwd)");
  c3_pack_kernel<<<grid_for(static_cast<long>(n) * h * w), kBlock, 0, S_(s)>>>(x, static_cast<uint4*>(xp), n, h * w);
  return sfk_check_launch("c3_pack");
}

int sfk_c3_unpack(const void* gp, float* gx, int n, int h, int w, sfk_stream_t s) {
  SFK_REQUIRE(gp && gx && sfk_aligned16(gp) && n >= 1 && h >= 1 && w >= 1, SFK_E_ARG, "c3_unpack: bad args");
  SFK_REQUIRE(!sfk_act_f32(), SFK_E_ARG, "c3_unpack: bf16 storage only");
  c3_unpack_kernel<<<grid_for(static_cast<long>(n) * h * w), kBlock, 0, S_(s)>>>(static_cast<const uint4*>(gp), gx, n, h * w);
  return sfk_check_launch("c3_unpack");
}

int sfk_avgpool_affine_fwd(const float* x, float* y, int n_planes, int h, int w, int k, float a, float b, sfk_stream_t s) {
  SFK_REQUIRE(x && y && k >= 1 && h % k == 0 && w % k == 0, SFK_E_ARG, "avgpool_affine: bad args");
  avgpool_affine_kernel<<<grid_for(static_cast<long>(n_planes) * (h / k) * (w / k)), kBlock, 0, S_(s)>>>(x, y, n_planes, h, w, k, a, b);
  return sfk_check_launch("avgpool_affine");
}

int sfk_maxpool2_fwd(const void* x, void* y, int n, int h, int w, int c, sfk_stream_t s) {
  SFK_REQUIRE(x && y && c % 8 == 0, SFK_E_ARG, "maxpool2_fwd: bad args");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      maxpool2_fwd_kernel<T><<<grid_for(static_cast<long>(n) * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8)), kBlock, 0, S_(s)>>>(
      static_cast<const T*>(x), static_cast<T*>(y), n, h, w, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("maxpool2_fwd");
}

int sfk_maxpool2_bwd(const void* x, const void* y, const void* gy, void* gx, const void* tap_ref, float tap_coef, int relu_mask, int n, int h,
                     int w, int c, sfk_stream_t s) {
  (void)y;
  SFK_REQUIRE(x && gy && gx && c % 8 == 0, SFK_E_ARG, "maxpool2_bwd: bad args");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      maxpool2_bwd_kernel<T><<<grid_for(static_cast<long>(n) * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8)), kBlock, 0, S_(s)>>>(
      static_cast<const T*>(x), static_cast<const T*>(gy), static_cast<T*>(gx), static_cast<const T*>(tap_ref), tap_coef, relu_mask,
      n, h, w, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("maxpool2_bwd");
}

int sfk_gap_fwd(const void* x, float* y, int n, int hw, int c, sfk_stream_t s) {
  SFK_REQUIRE(x && y && c % 8 == 0, SFK_E_ARG, "gap_fwd: bad args");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      gap_fwd_kernel<T><<<dim3(c / 8, n), kBlock, 0, S_(s)>>>(static_cast<const T*>(x), y, hw, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("gap_fwd");
}

int sfk_gap_bwd(const void* x, const float* gy, void* gx, int n, int hw, int c, sfk_stream_t s) {
  SFK_REQUIRE(x && gy && gx && c % 8 == 0, SFK_E_ARG, "gap_bwd: bad args");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      gap_bwd_kernel<T><<<grid_for(static_cast<long>(n) * hw * (c / 8)), kBlock, 0, S_(s)>>>(static_cast<const T*>(x), gy, static_cast<T*>(gx), n, hw, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("gap_bwd");
}

int sfk_mse_tap(const void* f, const void* ref, void* g, float* loss, float coef_loss, float coef_grad, int accumulate, int relu_mask, int n,
                long per_sample, sfk_stream_t s) {
  SFK_REQUIRE(f && ref && per_sample % 8 == 0, SFK_E_ARG, "mse_tap: bad args");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      mse_tap_kernel<T><<<dim3(per_sample_blocks(per_sample / 8, n), n), kBlock, 0, S_(s)>>>(static_cast<const T*>(f), static_cast<const T*>(ref),
                                                                                      static_cast<T*>(g), loss, coef_loss, coef_grad, accumulate,
                                                                                      relu_mask, per_sample);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("mse_tap");
}

int sfk_mse_f32(const float* a, const float* b, float* g, float* loss, float coef_loss, float coef_grad, int accumulate, int n, long per_sample,
                sfk_stream_t s) {
  SFK_REQUIRE(a && b, SFK_E_ARG, "mse_f32: null");
  mse_f32_kernel<<<dim3(per_sample_blocks(per_sample, n), n), kBlock, 0, S_(s)>>>(a, b, g, loss, coef_loss, coef_grad, accumulate, per_sample);
  return sfk_check_launch("mse_f32");
}

int sfk_image_loss_grad(const float* img, const float* ref, const float* gpool, float* g, float* loss, float coef_loss, float coef_grad, int n,
                        int size, int k, sfk_stream_t s) {
  SFK_REQUIRE(img && ref && g && size % k == 0 && size % 4 == 0, SFK_E_ARG, "image_loss_grad: bad args");
  image_loss_grad_kernel<<<dim3(per_sample_blocks(3L * size * size / 4, n), n), kBlock, 0, S_(s)>>>(img, ref, gpool, g, loss, coef_loss, coef_grad, size, k);
  return sfk_check_launch("image_loss_grad");
}

int sfk_style_affine_fwd(const float* w, const float* A, const float* bias, const int32_t* row_widx, float* sdst, int n, int n_latent,
                         int style_dim, int s_dim, float scale, sfk_stream_t st) {
  SFK_REQUIRE(w && A && bias && row_widx && sdst, SFK_E_ARG, "style_affine_fwd: null");
  style_affine_fwd_kernel<<<grid_for(static_cast<long>(s_dim) * 32), kBlock, 0, S_(st)>>>(w, A, bias, row_widx, sdst, n, n_latent, style_dim, s_dim, scale);
  return sfk_check_launch("style_affine_fwd");
}

int sfk_style_affine_bwd(const float* gs, const float* A, const int32_t* layer_row_start, const int32_t* layer_widx, int n_layers, float* gw, int n,
                         int n_latent, int style_dim, int s_dim, float scale, sfk_stream_t st) {
  SFK_REQUIRE(gs && A && layer_row_start && layer_widx && gw, SFK_E_ARG, "style_affine_bwd: null");
  zero_kernel<<<grid_for(static_cast<long>(n) * n_latent * style_dim), kBlock, 0, S_(st)>>>(gw, static_cast<long>(n) * n_latent * style_dim);
  style_affine_bwd_kernel<<<dim3(n_layers, 16, (n + 7) / 8), kBlock, 0, S_(st)>>>(gs, A, layer_row_start, layer_widx, gw, n, n_latent, style_dim, s_dim,
                                                                                   scale);
  return sfk_check_launch("style_affine_bwd");
}

int sfk_demod_fwd(const float* sv, int s_stride, const float* Q, float* d, int n, int cin, int cout, sfk_stream_t st) {
  SFK_REQUIRE(sv && Q && d, SFK_E_ARG, "demod_fwd: null");
  demod_fwd_kernel<<<grid_for(static_cast<long>(n) * cout * 32), kBlock, 0, S_(st)>>>(sv, s_stride, Q, d, n, cin, cout);
  return sfk_check_launch("demod_fwd");
}

int sfk_demod_bwd(const float* sv, int s_stride, const float* Q, const float* d, const float* gdacc, float* gs, int gs_stride, int n, int cin,
                  int cout, sfk_stream_t st) {
  SFK_REQUIRE(sv && Q && d && gdacc && gs, SFK_E_ARG, "demod_bwd: null");
  demod_bwd_kernel<<<dim3((cin + 31) / 32, n), 256, 0, S_(st)>>>(sv, s_stride, Q, d, gdacc, gs, gs_stride, n, cin, cout);
  return sfk_check_launch("demod_bwd");
}

int sfk_demod_fwd_batched(const float* sv, int s_stride, const float* q_cat, float* d_cat, const long long* tab, int n_layers, int n,
                          int max_cout, sfk_stream_t st) {
  SFK_REQUIRE(sv && q_cat && d_cat && tab && n_layers > 0 && max_cout > 0, SFK_E_ARG, "demod_fwd_batched: bad args");
  demod_fwd_batched_kernel<<<dim3(grid_for(static_cast<long>(n) * max_cout * 32, 2), n_layers), kBlock, 0, S_(st)>>>(sv, s_stride, q_cat, d_cat, tab, n);
  return sfk_check_launch("demod_fwd_batched");
}

int sfk_modulate_weights_batched(const float* wbase_cat, const float* sv, int s_stride, void* wmod_cat, const float* d_cat, const long long* tab,
                                 int n_layers, int n, sfk_stream_t st) {
  SFK_REQUIRE(wbase_cat && sv && wmod_cat && d_cat && tab && n_layers > 0 && s_stride % 4 == 0, SFK_E_ARG, "modulate_weights_batched: bad args");
  SFK_REQUIRE(sfk_aligned16(wbase_cat) && sfk_aligned16(sv) && sfk_aligned16(wmod_cat), SFK_E_ALIGN, "modulate_weights_batched: alignment");
  const unsigned gx = static_cast<unsigned>((sfk_num_sms() * 8 + n_layers - 1) / n_layers);
  if (sfk_act_f32())
    modulate_weights_batched_kernel<float><<<dim3(gx, n_layers), kBlock, 0, S_(st)>>>(wbase_cat, sv, s_stride, static_cast<float*>(wmod_cat), d_cat, tab, n);
  else
    modulate_weights_batched_kernel<bf16><<<dim3(gx, n_layers), kBlock, 0, S_(st)>>>(wbase_cat, sv, s_stride, static_cast<bf16*>(wmod_cat), d_cat, tab, n);
  return sfk_check_launch("modulate_weights_batched");
}

int sfk_demod_bwd_batched(const float* sv, int s_stride, const float* q_cat, const float* d_cat, const float* gd_cat, float* gs, int gs_stride,
                          const long long* tab, int n_layers, int n, int max_cin, sfk_stream_t st) {
  SFK_REQUIRE(sv && q_cat && d_cat && gd_cat && gs && tab && n_layers > 0 && max_cin > 0, SFK_E_ARG, "demod_bwd_batched: bad args");
  demod_bwd_batched_kernel<<<dim3((max_cin + 31) / 32, n, n_layers), 256, 0, S_(st)>>>(sv, s_stride, q_cat, d_cat, gd_cat, gs, gs_stride, tab);
  return sfk_check_launch("demod_bwd_batched");
}

int sfk_modulate_weights(const float* wbase, const float* sv, int s_stride, void* wmod, int n, int taps, int cout, int cin, const float* d,
                         int d_cols, sfk_stream_t st) {
  SFK_REQUIRE(wbase && sv && wmod && cin % 8 == 0 && s_stride % 4 == 0, SFK_E_ARG, "modulate_weights: bad args");
  SFK_REQUIRE(d == nullptr || (d_cols > 0 && cout % d_cols == 0), SFK_E_ARG, "modulate_weights: d_cols must divide cout");
  SFK_REQUIRE(sfk_aligned16(wbase) && sfk_aligned16(sv) && sfk_aligned16(wmod), SFK_E_ALIGN, "modulate_weights: alignment");
  const long rows = static_cast<long>(taps) * cout;
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      modulate_weights_kernel<T><<<grid_for(static_cast<long>(n) * rows * (cin / 8)), kBlock, 0, S_(st)>>>(wbase, sv, s_stride, static_cast<T*>(wmod), n, rows, cin, d, cout, d_cols > 0 ? d_cols : cout);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("modulate_weights");
}

static int blur_launch(bool bwd, const void* a0, const void* a1, void* dst, const float* d, const float* noise, float noise_w, const float* bias,
                       float* gdacc, const float* s_in, float* gs_in, int in_stride, int n, int h, int w, int c, sfk_stream_t st) {
  {
    const int rc = sfk_blur_stream_launch(bwd, a0, a1, dst, d, noise, noise_w, bias, gdacc, s_in, gs_in, in_stride, n, h, w, c, S_(st));
    if (rc != -1000) return rc;
  }
  if (sfk_act_f32()) {
    if (bwd)
      blur_simple_bwd_kernel<float><<<dim3(per_sample_blocks(static_cast<long>(2 * h + 2) * (2 * w + 2) * (c / 8), n), n), kBlock, 0, S_(st)>>>(
          static_cast<const float*>(a0), static_cast<const float*>(a1), static_cast<float*>(dst), d, noise, noise_w, bias, gdacc, s_in, gs_in, in_stride, h, w, c);
    else
      blur_simple_fwd_kernel<float><<<dim3(per_sample_blocks(4L * h * w * (c / 8), n), n), kBlock, 0, S_(st)>>>(
          static_cast<const float*>(a0), static_cast<float*>(dst), d, noise, noise_w, bias, h, w, c);
    return sfk_check_launch(bwd ? "blur_act_bwd(f32)" : "blur_act_fwd(f32)");
  }
  const int vecs = c / 8;
  const BlurGeom G = blur_geom(vecs);
  SFK_REQUIRE(vecs % G.CV == 0 && (G.CV == 1 || G.CV == 2 || G.CV == 4), SFK_E_SHAPE, "blur: channel count must be 8, 16 or a multiple of 32");
  const size_t smem = static_cast<size_t>(19) * G.pitch * G.CV * 16;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(blur_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(blur_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  const int rows = bwd ? 2 * h + 2 : 2 * h, cols = bwd ? 2 * w + 2 : 2 * w;
  const long tiles = static_cast<long>((rows + 15) / 16) * ((cols + G.TC - 1) / G.TC);
  const int cgs = vecs / G.CV;
  long bx = tiles;
  const long cap = (static_cast<long>(sfk_num_sms()) * 4 + n * cgs - 1) / (n * cgs);
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(static_cast<unsigned>(bx), n, cgs);
  if (bwd)
    blur_tile_kernel<true><<<grid, 256, smem, S_(st)>>>(static_cast<const bf16*>(a0), static_cast<const bf16*>(a1), static_cast<bf16*>(dst), d, noise,
                                                        noise_w, bias, gdacc, s_in, gs_in, in_stride, h, w, c);
  else
    blur_tile_kernel<false><<<grid, 256, smem, S_(st)>>>(static_cast<const bf16*>(a0), nullptr, static_cast<bf16*>(dst), d, noise, noise_w, bias,
                                                         nullptr, nullptr, nullptr, 0, h, w, c);
  return sfk_check_launch(bwd ? "blur_act_bwd" : "blur_act_fwd");
}

int sfk_blur_act_fwd(const void* T, void* out, const float* d, const float* noise, float noise_w, const float* bias, int n, int h, int w, int c,
                     sfk_stream_t st) {
  SFK_REQUIRE(T && out && d && bias && c % 8 == 0, SFK_E_ARG, "blur_act_fwd: bad args");
  return blur_launch(false, T, nullptr, out, d, noise, noise_w, bias, nullptr, nullptr, nullptr, 0, n, h, w, c, st);
}

int sfk_blur_act_bwd(const void* out, const void* gout, void* gT, const float* d, const float* noise, float noise_w, const float* bias,
                     float* gdacc, const float* s_in, float* gs_in, int vec_stride, int n, int h, int w, int c, sfk_stream_t st) {
  SFK_REQUIRE(out && gout && gT && d && bias && gdacc && c % 8 == 0 && (gs_in == nullptr || s_in != nullptr), SFK_E_ARG, "blur_act_bwd: bad args");
  return blur_launch(true, out, gout, gT, d, noise, noise_w, bias, gdacc, s_in, gs_in, vec_stride, n, h, w, c, st);
}

int sfk_act_bwd(const void* out, const void* gout, void* gz, const float* d, const float* noise, float noise_w, const float* bias, float* gdacc,
                const float* s_in, float* gs_in, int vec_stride, int n, int h, int w, int c, sfk_stream_t st) {
  SFK_REQUIRE(out && gout && gz && d && bias && gdacc && c % 8 == 0 && (c / 8) <= kBlock && kBlock % (c / 8) == 0, SFK_E_ARG, "act_bwd: bad args");
  {
    const int rc = sfk_act_stream_launch(out, gout, gz, d, noise, noise_w, bias, gdacc, nullptr, nullptr, 0, nullptr, nullptr, 0, s_in, gs_in,
                                         vec_stride, vec_stride, n, h * w, c, S_(st));
    if (rc != -1000) return rc;
  }
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      act_bwd_kernel<T><<<dim3(per_sample_blocks(static_cast<long>(h) * w * (c / 8), n), n), kBlock, c * sizeof(float), S_(st)>>>(
      static_cast<const T*>(out), static_cast<const T*>(gout), static_cast<T*>(gz), d, noise, noise_w, bias, gdacc, s_in, gs_in, vec_stride, h * w, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("act_bwd");
}

int sfk_torgb_fwd(const void* x, const float* wrgb, const float* sv, int s_stride, const float* bias, const float* skip, float* rgb, int n, int h,
                  int w, int c, sfk_stream_t st) {
  SFK_REQUIRE(x && wrgb && sv && bias && rgb && c % 8 == 0, SFK_E_ARG, "torgb_fwd: bad args");
  {
    const int rc = sfk_torgb_stream_launch(x, wrgb, sv, s_stride, bias, skip, rgb, n, h, w, c, S_(st));
    if (rc != -1000) return rc;
  }
  int lp = 1;
  while (lp < 32 && lp * 16 <= c / 8) lp *= 2;   // C<=128: one thread per pixel (full 64-256 B reads, coalesced planar stores)
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      torgb_fwd_kernel<T><<<dim3(per_sample_blocks(static_cast<long>(h) * w * lp, n), n), kBlock, 3 * c * sizeof(float), S_(st)>>>(
      static_cast<const T*>(x), wrgb, sv, s_stride, bias, skip, rgb, h, w, c, lp);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("torgb_fwd");
}

int sfk_torgb_bwd(const void* x, const float* wrgb, const float* sv, int s_stride, const float* grgb, void* gx, float* gs, int gs_stride, int n,
                  int h, int w, int c, sfk_stream_t st) {
  SFK_REQUIRE(x && wrgb && sv && grgb && gx && gs && c % 8 == 0 && (c / 8) <= kBlock && kBlock % (c / 8) == 0, SFK_E_ARG, "torgb_bwd: bad args");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      torgb_bwd_kernel<T><<<dim3(per_sample_blocks(static_cast<long>(h) * w * (c / 8), n), n), kBlock, 4 * c * sizeof(float), S_(st)>>>(
      static_cast<const T*>(x), wrgb, sv, s_stride, grgb, static_cast<T*>(gx), gs, gs_stride, h * w, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("torgb_bwd");
}

int sfk_act_torgb_bwd(const void* out, const void* gin, void* gz, const float* d, const float* noise, float noise_w, const float* bias,
                      float* gdacc, const float* wrgb, const float* sv, int s_stride, const float* grgb, float* gs, int gs_stride,
                      const float* s_in, float* gs_in, int n, int h, int w, int c, sfk_stream_t st) {
  SFK_REQUIRE(out && gz && d && bias && gdacc && wrgb && sv && grgb && gs && c % 8 == 0 && (c / 8) <= kBlock && kBlock % (c / 8) == 0, SFK_E_ARG,
              "act_torgb_bwd: bad args");
  {
    const int rc = sfk_act_stream_launch(out, gin, gz, d, noise, noise_w, bias, gdacc, wrgb, sv, s_stride, grgb, gs, gs_stride, s_in, gs_in,
                                         s_stride, gs_stride, n, h * w, c, S_(st));
    if (rc != -1000) return rc;
  }
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      act_torgb_bwd_kernel<T><<<dim3(per_sample_blocks(static_cast<long>(h) * w * (c / 8), n), n), kBlock, 4 * c * sizeof(float), S_(st)>>>(
          static_cast<const T*>(out), static_cast<const T*>(gin), static_cast<T*>(gz), d, noise, noise_w, bias, gdacc, wrgb, sv, s_stride, grgb, gs,
          gs_stride, s_in, gs_in, h * w, c);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("act_torgb_bwd");
}

int sfk_rgb_down(const float* g, float* gskip, int planes, int h, int w, sfk_stream_t st) {
  SFK_REQUIRE(g && gskip && h % 2 == 0 && w % 2 == 0, SFK_E_ARG, "rgb_down: bad args");
  if (w % 4 == 0 && sfk_aligned16(g) && (reinterpret_cast<uintptr_t>(gskip) & 7) == 0)
    rgb_down2_kernel<<<grid_for(static_cast<long>(planes) * (h / 2) * (w / 4)), kBlock, 0, S_(st)>>>(g, gskip, planes, h, w);
  else
    rgb_down_kernel<<<grid_for(static_cast<long>(planes) * (h / 2) * (w / 2)), kBlock, 0, S_(st)>>>(g, gskip, planes, h, w);
  return sfk_check_launch("rgb_down");
}

int sfk_linear_fwd(const float* x, const float* W, const float* bias, float* y, int n, int in, int out, sfk_stream_t st) {
  SFK_REQUIRE(x && W && y, SFK_E_ARG, "linear_fwd: null");
  linear_fwd_kernel<<<grid_for(static_cast<long>(out) * 32), kBlock, 0, S_(st)>>>(x, W, bias, y, n, in, out);
  return sfk_check_launch("linear_fwd");
}

int sfk_linear_bwd(const float* gy, const float* W, float* gx, int n, int in, int out, sfk_stream_t st) {
  SFK_REQUIRE(gy && W && gx, SFK_E_ARG, "linear_bwd: null");
  zero_kernel<<<grid_for(static_cast<long>(n) * in), kBlock, 0, S_(st)>>>(gx, static_cast<long>(n) * in);
  linear_bwd_kernel<<<dim3((in + 127) / 128, (out + kLinSlice - 1) / kLinSlice), 128, 0, S_(st)>>>(gy, W, gx, n, in, out);
  return sfk_check_launch("linear_bwd");
}

int sfk_fuse_spatial_fwd(const float* sa, const float* sb, const float* al, const float* be, const float* c, float* s, int n, int dim, sfk_stream_t st) {
  SFK_REQUIRE(sa && sb && al && be && c && s, SFK_E_ARG, "fuse_spatial_fwd: null");
  fuse_spatial_fwd_kernel<<<grid_for(static_cast<long>(n) * dim), kBlock, 0, S_(st)>>>(sa, sb, al, be, c, s, n, dim);
  return sfk_check_launch("fuse_spatial_fwd");
}

int sfk_fuse_spatial_bwd(const float* sa, const float* sb, const float* al, const float* be, const float* c, const float* gs, float* gsa, float* gsb,
                         int n, int dim, sfk_stream_t st) {
  SFK_REQUIRE(sa && sb && al && be && c && gs && gsa && gsb, SFK_E_ARG, "fuse_spatial_bwd: null");
  fuse_spatial_bwd_kernel<<<grid_for(static_cast<long>(n) * dim), kBlock, 0, S_(st)>>>(sa, sb, al, be, c, gs, gsa, gsb, n, dim);
  return sfk_check_launch("fuse_spatial_bwd");
}

int sfk_axpby(const float* x, const float* y, float* out, float a, float b, long n, sfk_stream_t st) {
  SFK_REQUIRE(x && out, SFK_E_ARG, "axpby: null");
  axpby_kernel<<<grid_for(n), kBlock, 0, S_(st)>>>(x, y, out, a, b, n);
  return sfk_check_launch("axpby");
}

int sfk_nchw_to_nhwc_bf16(const float* x, void* y, int n, int c, int h, int w, sfk_stream_t st) {
  SFK_REQUIRE(x && y, SFK_E_ARG, "nchw_to_nhwc: null");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      nchw_to_nhwc_kernel<T><<<grid_for(static_cast<long>(n) * c * h * w), kBlock, 0, S_(st)>>>(x, static_cast<T*>(y), n, c, h, w);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("nchw_to_nhwc");
}

int sfk_nhwc_bf16_to_nchw(const void* x, float* y, int n, int c, int h, int w, sfk_stream_t st) {
  SFK_REQUIRE(x && y, SFK_E_ARG, "nhwc_to_nchw: null");
  {
    auto run = [&](auto tag) {
      using T = decltype(tag);
      nhwc_to_nchw_kernel<T><<<grid_for(static_cast<long>(n) * c * h * w), kBlock, 0, S_(st)>>>(static_cast<const T*>(x), y, n, c, h, w);
    };
    if (sfk_act_f32()) run(float{}); else run(bf16{});
  }
  return sfk_check_launch("nhwc_to_nchw");
}

int sfk_attack_update_linf(float* x, const float* x0, const float* gpool, float alpha, float eps, float dir, float lo, float hi, float* stats,
                           int n, int size, int k, sfk_stream_t st) {
  SFK_REQUIRE(x && x0 && gpool && size % k == 0 && size % 4 == 0, SFK_E_ARG, "attack_update_linf: bad args");
  update_linf_kernel<<<dim3(per_sample_blocks(3L * size * size / 4, n), n), kBlock, 0, S_(st)>>>(x, x0, gpool, alpha, eps, dir, lo, hi, stats, size, k);
  return sfk_check_launch("attack_update_linf");
}

int sfk_attack_random_start(float* x, const float* x0, float eps, float lo, float hi, unsigned long long seed, long count, sfk_stream_t st) {
  SFK_REQUIRE(x && x0 && count > 0 && count % 4 == 0 && sfk_aligned16(x) && sfk_aligned16(x0), SFK_E_ARG, "attack_random_start: bad args");
  random_start_kernel<<<grid_for(count / 4), kBlock, 0, S_(st)>>>(reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(x0), eps, lo, hi,
                                                                 static_cast<uint64_t>(seed), count / 4);
  return sfk_check_launch("attack_random_start");
}

int sfk_attack_update_patch(float* x, const float* x0, float* patch, const float* mask, const float* gpool, float lr, float dir, int use_sign,
                            const float* lo, const float* hi, float gscale, float* stats, int n, int size, int k, sfk_stream_t st) {
  SFK_REQUIRE(x && x0 && patch && mask && gpool && lo && hi && size % k == 0, SFK_E_ARG, "attack_update_patch: bad args");
  update_patch_kernel<<<dim3(per_sample_blocks(3L * size * size, n), n), kBlock, 0, S_(st)>>>(x, x0, patch, mask, gpool, lr, dir, use_sign, lo, hi, gscale,
                                                                                             stats, size, k);
  return sfk_check_launch("attack_update_patch");
}

int sfk_attack_update_adam(float* x, const float* gpool, const float* gfull, float gfull_scale, float* m, float* v, float lr, float b1, float b2,
                           float eps, int t, float gscale, int n, int size, int k, sfk_stream_t st) {
  SFK_REQUIRE(x && gpool && m && v && t >= 1 && size % k == 0, SFK_E_ARG, "attack_update_adam: bad args");
  const float bc1 = 1.f - powf(b1, static_cast<float>(t)), bc2 = 1.f - powf(b2, static_cast<float>(t));
  update_adam_kernel<<<dim3(per_sample_blocks(3L * size * size, n), n), kBlock, 0, S_(st)>>>(x, gpool, gfull, gfull_scale, m, v, lr, b1, b2, eps, bc1, bc2, gscale, size, k);
  return sfk_check_launch("attack_update_adam");
}

int sfk_attack_update_l2(float* x, const float* x0, const float* gpool, float* norms, float* dn, float alpha, float eps, float dir, float lo,
                         float hi, int phase, int n, int size, int k, sfk_stream_t st) {
  SFK_REQUIRE(x && x0 && gpool && norms && dn && phase >= 0 && phase <= 2 && size % k == 0, SFK_E_ARG, "attack_update_l2: bad args");
  update_l2_kernel<<<dim3(per_sample_blocks(3L * size * size, n), n), kBlock, 0, S_(st)>>>(x, x0, gpool, norms, dn, alpha, eps, dir, lo, hi, phase, size, k);
  return sfk_check_launch("attack_update_l2");
}

int sfk_minmax_per_sample(const float* x, float* lo, float* hi, int n, long per_sample, sfk_stream_t st) {
  SFK_REQUIRE(x && lo && hi, SFK_E_ARG, "minmax: null");
  minmax_kernel<<<n, 1024, 0, S_(st)>>>(x, lo, hi, per_sample);
  return sfk_check_launch("minmax");
}

int sfk_patch_grad_reduce(const float* gpool, const float* mask, float* gsum, float gscale, int n, int size, int k, sfk_stream_t st) {
  SFK_REQUIRE(gpool && mask && gsum, SFK_E_ARG, "patch_grad_reduce: null");
  SFK_REQUIRE(n >= 1 && size >= 1 && k >= 1 && size % k == 0, SFK_E_SHAPE, "patch_grad_reduce: bad shape");
  patch_grad_reduce_kernel<<<grid_for(3L * size * size), kBlock, 0, S_(st)>>>(gpool, mask, gsum, gscale, n, size, k);
  return sfk_check_launch("patch_grad_reduce");
}

int sfk_patch_apply_shared(float* x, const float* x0, const float* patch, const float* mask, const float* lo, const float* hi, int n,
                           int size, sfk_stream_t st) {
  SFK_REQUIRE(x && x0 && patch && mask && lo && hi, SFK_E_ARG, "patch_apply_shared: null");
  dim3 grid(per_sample_blocks(3L * size * size, n), n);
  patch_apply_shared_kernel<<<grid, kBlock, 0, S_(st)>>>(x, x0, patch, mask, lo, hi, size);
  return sfk_check_launch("patch_apply_shared");
}

int sfk_ssim_gray7(const float* a, const float* b, float* out, int n, int h, int w, float data_range, sfk_stream_t st) {
  SFK_REQUIRE(a && b && out, SFK_E_ARG, "ssim: null");
  SFK_REQUIRE(n >= 1 && h >= kSsimW && w >= kSsimW, SFK_E_SHAPE, "ssim: images must be at least 7x7");
  const int vh = h - 2 * kSsimR, vw = w - 2 * kSsimR;     // window centres
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * n, S_(st));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid((vw + kSsimT - 1) / kSsimT, (vh + kSsimT - 1) / kSsimT, n);
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  ssim_gray7_kernel<<<grid, kSsimT * kSsimT, 0, S_(st)>>>(a, b, out, h, w, c1, c2, 1.f / (static_cast<float>(vh) * vw));
  return sfk_check_launch("ssim_gray7");
}

}  // extern "C"
