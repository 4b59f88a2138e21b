/* sfk.h -- C ABI of libsfattack.so: hand-written sm_100a kernels for the attack hot path.
 *
 * The reference (Wu-sm/Adversarial-Attacks-on-GAN-based-Image-Fusion) has NO native interface: its
 * hot loops call PyTorch modules.  Each entry point below therefore cites the reference
 * Python call it stands under.  All pointers are DEVICE pointers owned by the caller, 16-byte aligned;
 * nothing here allocates, synchronises or touches the default stream; every call enqueues on `stream`.
 * Return: 0 ok, >0 cudaError_t, <0 argument error (SFK_E_*).
 *
 * Layouts:  activations NHWC bf16 [N][H][W][C];  images NCHW fp32 [N][3][H][W];
 *           phase-planar (stride-2 transposed conv intermediates) [N][4][H+1][W+1][C] bf16 with
 *           plane p=2a+b holding T[2m+a][2n+b];  per-sample vectors fp32 [N][C];
 *           conv weights bf16 [S][taps][Cout][Cin] (S=1 shared, S=N per-sample modulated).
 */
#ifndef SFK_H
#define SFK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sfk_stream_t; /* cudaStream_t */

#define SFK_E_ARG (-1)
#define SFK_E_SHAPE (-2)
#define SFK_E_DRIVER (-3)
#define SFK_E_ALIGN (-4)

int sfk_version(void);
/* Activation storage of every entry point: 0 = bf16 (default, product path), 1 = fp32 (parity mode: the same schedules at fp32
 * storage; the conv runs the same tcgen05 pipeline with kind::tf32 MMAs, see sfk_set_conv_math).  Process-global; set before
 * allocating. */
int sfk_set_activation_dtype(int f32);
int sfk_get_activation_dtype(void);
const char* sfk_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
 * One description covers every dense contraction on the path:
 *   VGG / encoder conv3x3+bias+ReLU fwd and dgrad     (code/vgg.py:45-63 and its autograd)
 *   StyleGAN2 ModulatedConv2d fwd and dgrad            (decoder(...) call, attack_main2.py:619-621)
 *   stride-2 transposed conv as 4 phase accumulators   (same call, upsample=True layers)
 * GEMM view: M = output positions (tile = TH x TW = 128), N = block_n output channels per
 * accumulator, K = taps x Cin.  A k-step loads the activation box shifted by (dy,dx) from plane
 * `plane` and the weight rows [brow + n0, +block_n) and accumulates into accumulator `acc`. */
#define SFK_MAX_TAPS 16
typedef struct {
  int32_t dy, dx;    /* input shift */
  int32_t plane;     /* input phase plane (0 for plain NHWC) */
  int32_t acc;       /* accumulator (output plane) */
  int32_t brow;      /* first weight row of this tap: tap_index * Cout */
} sfk_tap;

enum {
  SFK_EP_DSCALE = 1,     /* acc *= dscale[n][col]               (demodulation) */
  SFK_EP_NOISE = 2,      /* acc += noise_w * noise[h][w]        (NoiseInjection) */
  SFK_EP_BIAS = 4,       /* acc += bias[col] */
  SFK_EP_RELU = 8,
  SFK_EP_LRELU = 16,     /* leaky_relu(.,0.2)*sqrt(2)           (FusedLeakyReLU) */
  SFK_EP_XMASK = 32,     /* acc *= (xin > 0)                    (ReLU backward, fused into dgrad) */
  SFK_EP_GSDOT = 64,     /* gs[n][col] += sum_hw xin*acc        (style gradient, fused into dgrad) */
  SFK_EP_COLSCALE = 128, /* acc *= colscale[n][col]             (gx = s * gx~) */
  SFK_EP_ACCUM = 256,    /* out += acc                          (second consumer of an activation) */
  SFK_EP_LRELU_RAW = 512 /* max(acc, 0.2*acc): FusedLeakyReLU whose gain sqrt(2) the caller has folded into the weights, bias
                            and noise strength (lrelu(x)*g == lrelu(g*x) for g > 0): one instruction less per element */
};

typedef struct {
  /* A operand: activations */
  const void* a;             /* bf16 */
  int32_t n_img, a_h, a_w, a_c, a_planes; /* A tensor dims: [n_img][a_planes][a_h][a_w][a_c] */
  /* B operand: weights [b_samples][b_rows][a_c] bf16, K(=a_c)-major */
  const void* b;
  int32_t b_samples, b_rows;
  /* iteration space / output: [n_img][num_acc][out_h][out_w][out_c] bf16 */
  void* out;
  int32_t out_h, out_w, out_c;
  int32_t num_acc, block_n;  /* num_acc*block_n <= 512, block_n % 16 == 0, block_n <= 256 */
  int32_t num_taps;
  sfk_tap taps[SFK_MAX_TAPS];
  /* epilogue */
  int32_t flags;
  const float* dscale;       /* [n_img][out_c] */
  const float* bias;         /* [out_c] */
  const float* noise;        /* [out_h][out_w] */
  float noise_w;
  const void* xin;           /* bf16 [n_img][out_h][out_w][out_c] */
  const float* colscale;     /* [n_img][out_c] */
  float* gs;                 /* [n_img][out_c] */
  int32_t vec_stride;        /* row stride (floats) of colscale and gs; 0 = out_c */
  int32_t* err;              /* device int, set non-zero on an internal timeout */
  int32_t stages;            /* 0 = auto */
  /* Resample fused into the conv (upsampling StyledConv = 4 phase convs over the INPUT grid):
   * out_d2s: the out_c = 4*Cq columns are the 4 output phases; column p*Cq+c of position (h,w) is stored to pixel
   *          (2h + p/2, 2w + p%2), channel c of an [n][2*out_h][2*out_w][Cq] tensor; dscale/bias are indexed by c, noise by the
   *          fine pixel.  (depth-to-space epilogue)
   * a_s2d:   A is physically [n][2*a_h][2*a_w][a_c/4]; its K index is p*Cq+c for fine pixel (2h+p/2, 2w+p%2) (space-to-depth
   *          view, used by the data gradient of the fused op). */
  int32_t out_d2s, a_s2d;
  /* scratch for the split-tf32 parity mode (fp32 storage, conv math 0/2): >= sfk_igemm_workspace_bytes(desc) bytes, 16-byte
   * aligned, private to this launch while it runs (launches on one stream may share it).  Unused (may be NULL) otherwise. */
  void* ws;
  size_t ws_bytes;
} sfk_igemm_desc;

/* ---- weight gradients: the second backward GEMM of the 3x3 convolutions (north_star: "forward and both backward GEMMs").
 * The reference leaves every parameter trainable (code/attack/attack_main2.py:301-304), so its autograd evaluates these on every
 * iteration of the loops at attack_main2.py:584-671 / adversarial_patch.py:94-160 although only the pixel gradient is used.
 *
 * sfk_conv3x3_wgrad: dw[s][tap][co][ci] += sum_{n in s} sum_{h,w} gz[n][h][w][co] * x[n][h+dy][w+dx][ci], tap = (dy+1)*3 + (dx+1),
 *   zero padding; x [n][h][w][cin], gz [n][h][w][cout] in the activation storage type, dw fp32 [S][9][cout][cin] with S = n
 *   (per_sample, ModulatedConv2d) or 1 (shared weights, VGG / encoder); dw is ACCUMULATED into (zero it first).  tcgen05 kernel
 *   (MN-major operands straight from the NHWC tensors, csrc/sfk_wgrad.cu) for bf16 storage, w >= 8, channels % 8 == 0; CUDA cores
 *   otherwise or when use_ref != 0.  err: device int raised on an internal pipeline timeout (may be NULL).
 * sfk_bias_grad: db[c] += sum_{n,h,w} gz (conv3x3+bias+ReLU; gz already carries the ReLU mask).
 * sfk_modconv_wgrad_finish: from the per-sample GEMM result G [n][9][cout][cin] (gz = d * dL/dy, x the unmodulated input) to the
 *   gradient of the shared base weight Wb = scale*W [9][cout][cin]:
 *   dWb = sum_n s[n][ci] * (G[n] - gdacc[n][co] * d[n][co]^2 * Wb * s[n][ci])   (second term only if demodulate). */
int sfk_conv3x3_wgrad(const void* x, const void* gz, float* dw, int n, int h, int w, int cin, int cout, int per_sample, int use_ref,
                      int32_t* err, sfk_stream_t stream);
int sfk_bias_grad(const void* gz, float* db, int n, int hw, int c, sfk_stream_t stream);
/* first conv (3 input channels, x fp32 NCHW as sfk_conv_c3_fwd takes it): dw [cout][3][3][3] (torch layout) += ... */
int sfk_conv_c3_wgrad(const float* x, const void* gz, float* dw, int n, int h, int w, int cout, sfk_stream_t stream);
int sfk_modconv_wgrad_finish(const float* G, const float* wb, const float* s, int s_stride, const float* d, const float* gdacc,
                             float* dwb, int n, int cout, int cin, int demodulate, sfk_stream_t stream);

/* One-shot launch: plans (tap grouping, shared-memory plan, tensor-map encoding) and launches.  Re-entrant: no static state. */
int sfk_igemm(const sfk_igemm_desc* d, sfk_stream_t stream);
/* Prepared launches: sfk_igemm_prepare does all the host-side work once and returns an immutable plan bound to the descriptor's
 * buffers, storage mode and conv math; sfk_igemm_run is then a single kernel launch and may be called concurrently from any
 * number of host threads / streams.  (The reference has no counterpart: torch caches cuDNN plans the same way.) */
typedef struct sfk_igemm_plan sfk_igemm_plan;
int sfk_igemm_prepare(const sfk_igemm_desc* d, sfk_igemm_plan** plan);
int sfk_igemm_run(const sfk_igemm_plan* plan, sfk_stream_t stream);
int sfk_igemm_destroy(sfk_igemm_plan* plan);
/* planner decisions of a prepared launch (tests, profiling): out16 = { two M tiles per stage, halo loads, resident weights,
 * depth-to-space out, space-to-depth in, passes, stages, block_n, compile-time variant, epilogue flags, grid.x, grid.y,
 * dynamic smem bytes, fp32 storage, CUDA-core kernel, accumulator stages * 16 + partial accumulators } */
int sfk_igemm_plan_info(const sfk_igemm_plan* plan, int32_t* out16);
/* Arithmetic of the tensor-core conv when the activation storage is fp32 (sfk_set_activation_dtype(1)).  In every tensor-core
 * mode the k-blocks rotate over up to 8 partial TMEM accumulators that the epilogue sums: the tensor core adds into its fp32
 * accumulator with truncation, which would otherwise cost ~1.2e-8 relative per accumulated element.
 *   0 / 2  split tf32: three kind::tf32 passes over hi/lo-split operands (A.hi*B.hi + A.lo*B.hi + A.hi*B.lo), products accurate to
 *          ~2^-21 relative -- the mode held to north_star's 1e-3 tolerance;   needs desc.ws
 *   1      plain kind::tf32 (10-bit mantissa operands, fp32 accumulate)
 *   3      CUDA-core kernel (sfk_igemm_ref), the third opinion
 * bf16 storage always uses kind::f16 bf16 MMAs.  Process-global; set before preparing plans. */
int sfk_set_conv_math(int mode);
int sfk_get_conv_math(void);
size_t sfk_igemm_workspace_bytes(const sfk_igemm_desc* d);
/* One-tile-per-CTA variant of the same contract (first implementation; kept for A/B timing and as a second cross-check). */
int sfk_igemm_v1(const sfk_igemm_desc* d, sfk_stream_t stream);
/* Profiling aid: with flag bit 16 set, sfk_igemm accumulates per-role wait/total cycles (producer, MMA issuer, epilogue);
 * this copies the 8 counters out (synchronising) and optionally resets them. */
int sfk_role_cycles(unsigned long long* out8, int reset);
/* Same contract on CUDA cores with plain loops: the on-device cross-check of the tensor-core path. */
int sfk_igemm_ref(const sfk_igemm_desc* d, sfk_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * First-layer 3x3 conv (Cin = 3, K = 27: bandwidth-bound, CUDA cores).  code/vgg.py:45 (conv1_1). */
int sfk_conv_c3_fwd(const float* x /*[N][3][H][W]*/, const float* w /*[Cout][3][3][3]*/, const float* bias,
                    void* out /*bf16 NHWC*/, int n, int h, int w_, int cout, int relu, sfk_stream_t s);
/* dL/dx of the same conv; g is already the PRE-activation gradient. */
int sfk_conv_c3_bwd(const void* g /*bf16 NHWC*/, const float* w, float* gx /*[N][3][H][W]*/,
                    int n, int h, int w_, int cout, sfk_stream_t s);

/* The same first conv on the tensor cores (bf16 storage): the 3-channel fp32 image becomes a 16-channel NHWC bf16 operand
 * [x.hi(3) | x.lo(3) | x.hi(3) | 0 x 7] (x = hi + lo, both bf16) that sfk_igemm multiplies with weights laid out
 * [W.hi(3) | W.hi(3) | W.lo(3) | 0 x 7] per tap (fp32-class products, one K = 16 MMA per tap); the data gradient comes back
 * from sfk_igemm as 16 bf16 channels [g.W.hi(3) | g.W.lo(3) | ...] that sfk_c3_unpack sums into gx [N][3][H][W] fp32.
 * code/vgg.py:45 (conv1_1) and the encoder's first conv. */
int sfk_c3_pack(const float* x /*[N][3][H][W]*/, void* xp /*bf16 [N][H][W][16]*/, int n, int h, int w, sfk_stream_t s);
int sfk_c3_unpack(const void* gp /*bf16 [N][H][W][16]*/, float* gx /*[N][3][H][W]*/, int n, int h, int w, sfk_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Pools.  F.avg_pool2d(img, k, k) (attack_main2.py:590-591,619-624) fused with y = a*pool(x)+b;
 * nn.MaxPool2d(2,2[,ceil_mode]) (code/vgg.py:14,18,24). */
int sfk_avgpool_affine_fwd(const float* x, float* y, int n_planes, int h, int w, int k, float a, float b, sfk_stream_t s);
int sfk_maxpool2_fwd(const void* x, void* y, int n, int h, int w, int c, sfk_stream_t s); /* out = ceil(h/2) */
/* gx = route(gy) [+ tap_coef*(x - tap_ref)] ; if relu_mask: gx *= (x > 0).  x = pool input (post-ReLU). */
int sfk_maxpool2_bwd(const void* x, const void* y, const void* gy, void* gx, const void* tap_ref, float tap_coef,
                     int relu_mask, int n, int h, int w, int c, sfk_stream_t s);
int sfk_gap_fwd(const void* x, float* y, int n, int hw, int c, sfk_stream_t s);       /* mean over H*W */
int sfk_gap_bwd(const void* x, const float* gy, void* gx, int n, int hw, int c, sfk_stream_t s); /* with ReLU mask */

/* ---------------------------------------------------------------------------------------------
 * MSE feature loss taps (nn.MSELoss(reduction='mean') per sample, attack_main2.py:605,626-645).
 *   loss[n] += coef_loss * sum((f-ref)^2);  g (=|+=) coef_grad*(f-ref) [* (f>0)] */
int sfk_mse_tap(const void* f, const void* ref, void* g, float* loss, float coef_loss, float coef_grad,
                int accumulate, int relu_mask, int n, long per_sample, sfk_stream_t s);
/* image term + avg-pool backward of the VGG input gradient:
 *   g[n][c][h][w] = coef_grad*(img-ref) + gpool[n][c][h/k][w/k]/k^2 ; loss[n] += coef_loss*sum((img-ref)^2) */
/* fp32 variant for latent codes (attack_main2.py:644-645 l_latent_target / l_latent_org) */
int sfk_mse_f32(const float* a, const float* b, float* g, float* loss, float coef_loss, float coef_grad, int accumulate,
                int n, long per_sample, sfk_stream_t s);
int sfk_image_loss_grad(const float* img, const float* ref, const float* gpool, float* g, float* loss,
                        float coef_loss, float coef_grad, int n, int size, int k, sfk_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * StyleGAN2 synthesis pieces (un-vendored stylefusion.sf_stylegan2). */
/* s[n][r] = bias[r] + scale * dot(w[n][row_widx[r]][:], A[r][:]) */
int sfk_style_affine_fwd(const float* w, const float* A, const float* bias, const int32_t* row_widx, float* s,
                         int n, int n_latent, int style_dim, int s_dim, float scale, sfk_stream_t st);
/* gw[n][l][k] = scale * sum_{r: row_widx[r]==l} gs[n][r]*A[r][k]   (rows of one layer are contiguous) */
int sfk_style_affine_bwd(const float* gs, const float* A, const int32_t* layer_row_start, const int32_t* layer_widx,
                         int n_layers, float* gw, int n, int n_latent, int style_dim, int s_dim, float scale,
                         sfk_stream_t st);
/* d[n][j] = rsqrt(sum_i s[n][i]^2 Q[j][i] + 1e-8) */
int sfk_demod_fwd(const float* s, int s_stride, const float* Q, float* d, int n, int cin, int cout, sfk_stream_t st);
/* gs[n][i] -= s[n][i] * sum_j gdacc[n][j] d[n][j]^2 Q[j][i] */
int sfk_demod_bwd(const float* s, int s_stride, const float* Q, const float* d, const float* gdacc, float* gs,
                  int gs_stride, int n, int cin, int cout, sfk_stream_t st);
/* wmod[n][t][j][i] = bf16(wbase[t][j][i] * s[n][i] * (d ? d[n][j % d_cols] : 1))   (wbase already carries 1/sqrt(cin*k*k)).
 * d (optional, [n][d_cols]) folds the demodulation of ModulatedConv2d into the weights so the conv epilogue has no
 * per-column scale; d_cols divides cout (the fused upsample conv stacks 4 phases of d_cols output channels in one tap). */
int sfk_modulate_weights(const float* wbase, const float* s, int s_stride, void* wmod, int n, int taps, int cout,
                         int cin, const float* d, int d_cols, sfk_stream_t st);
/* The three style-space kernels above for ALL modulated-conv layers of the generator in one launch each.  `tab` is a device table of
 * int64 [n_layers][SFK_STYLE_TAB_COLS]: { s_off, cin, cout, q_off, d_off, rows, wb_off, wm_off, d_cols, fold }.
 *   q_cat  : concatenated Q (cout x cin per layer, at q_off)          d_cat / gd_cat : [n][cout] per layer at d_off
 *   wbase_cat : [rows][cin] per layer at wb_off (rows = 9*cout, or 9*4*cout for a fused upsample conv, d_cols = cout)
 *   wmod_cat  : [n][rows][cin] per layer at n*wm_off;  fold != 0 multiplies the demodulation d into the weights, fold == 2
 *               also the activation gain sqrt(2) (for the SFK_EP_LRELU_RAW epilogue) */
#define SFK_STYLE_TAB_COLS 10
int sfk_demod_fwd_batched(const float* s, int s_stride, const float* q_cat, float* d_cat, const long long* tab, int n_layers,
                          int n, int max_cout, sfk_stream_t st);
int sfk_modulate_weights_batched(const float* wbase_cat, const float* s, int s_stride, void* wmod_cat, const float* d_cat,
                                 const long long* tab, int n_layers, int n, sfk_stream_t st);
int sfk_demod_bwd_batched(const float* s, int s_stride, const float* q_cat, const float* d_cat, const float* gd_cat, float* gs,
                          int gs_stride, const long long* tab, int n_layers, int n, int max_cin, sfk_stream_t st);
/* upfirdn2d([1,3,3,1] blur, pad (1,1)) of the phase-planar transposed-conv output, fused with
 * demod * . + noise + bias, leaky_relu*sqrt2.  T: [n][4][h+1][w+1][c] -> out [n][2h][2w][c]. */
int sfk_blur_act_fwd(const void* T, void* out, const float* d, const float* noise, float noise_w, const float* bias,
                     int n, int h, int w, int c, sfk_stream_t st);
/* backward of the above: gT (phase-planar) = blur^T(d * act'(out) * g); gdacc[n][c] += sum gy*y.
 * s_in == NULL: g = gout.  Otherwise gout is the plain (flags 0) data gradient gx~ of the conv that consumes `out`, and this
 * kernel finishes it as sfk_act_bwd does: g = s_in[n][c] * gout, gs_in[n][c] += sum_hw out * gout (rows of vec_stride floats). */
int sfk_blur_act_bwd(const void* out, const void* gout, void* gT, const float* d, const float* noise, float noise_w,
                     const float* bias, float* gdacc, const float* s_in, float* gs_in, int vec_stride, int n, int h, int w,
                     int c, sfk_stream_t st);
/* backward of FusedLeakyReLU+noise+demod for non-upsampling layers: gz = d*act'(out)*gout (in place ok) */
int sfk_act_bwd(const void* out, const void* gout, void* gz, const float* d, const float* noise, float noise_w,
                const float* bias, float* gdacc, const float* s_in, float* gs_in, int vec_stride, int n, int h, int w, int c,
                sfk_stream_t st);
/* ToRGB: rgb[n][c][h][w] = sum_i wrgb[c][i] s[n][i] x[n][h][w][i] + bias[c] + upsample2(skip) */
int sfk_torgb_fwd(const void* x, const float* wrgb, const float* s, int s_stride, const float* bias, const float* skip,
                  float* rgb, int n, int h, int w, int c, sfk_stream_t st);
/* gx[n][h][w][i] = s[n][i]*sum_c wrgb[c][i] grgb[n][c][h][w];  gs[n][i] += sum_hw x*gx~ */
int sfk_torgb_bwd(const void* x, const float* wrgb, const float* s, int s_stride, const float* grgb, void* gx,
                  float* gs, int gs_stride, int n, int h, int w, int c, sfk_stream_t st);
/* s_in / gs_in (both optional, in sfk_act_bwd and sfk_act_torgb_bwd): when the incoming gradient was written by a data-gradient
 * launch WITHOUT its style epilogue (flags 0), the consumer finishes it while it streams the two tensors anyway:
 *   gs_in[n][c] += sum_hw out * gin   (style gradient of the consuming conv; its input IS this layer's output)
 *   gin <- s_in[n][c] * gin           (modulation of that conv).  Strides: vec_stride (act_bwd) / s_stride, gs_stride (act_torgb_bwd). */
/* sfk_torgb_bwd + sfk_act_bwd of the conv feeding the ToRGB in ONE pass (StyledConv -> ToRGB):
 *   g = (gin ? gin : 0) + s_rgb * (wrgb^T grgb);  gs_rgb += sum_hw out * (wrgb^T grgb);  gz = d * g * act'(out);  gdacc += sum_hw gy*y
 * gin may be NULL (top resolution) and may alias gz. */
int sfk_act_torgb_bwd(const void* out, const void* gin, void* gz, const float* d, const float* noise, float noise_w,
                      const float* bias, float* gdacc, const float* wrgb, const float* s_rgb, int s_stride,
                      const float* grgb, float* gs_rgb, int gs_stride, const float* s_in, float* gs_in, int n, int h, int w,
                      int c, sfk_stream_t st);
/* transpose of the skip upsample: gskip = upfirdn2d(g, k*4, down=2, pad=(1,2)); planes = n*3 */
int sfk_rgb_down(const float* g, float* gskip, int planes, int h, int w, sfk_stream_t st);

/* ---------------------------------------------------------------------------------------------
 * small dense helpers (encoder head, fusion) */
int sfk_linear_fwd(const float* x, const float* W, const float* bias, float* y, int n, int in, int out, sfk_stream_t st);
int sfk_linear_bwd(const float* gy, const float* W, float* gx, int n, int in, int out, sfk_stream_t st);
/* spatial pair fusion stand-in: q = sigmoid(al*sa + be*sb + c); s = q*sa + (1-q)*sb */
int sfk_fuse_spatial_fwd(const float* sa, const float* sb, const float* al, const float* be, const float* c,
                         float* s, int n, int dim, sfk_stream_t st);
int sfk_fuse_spatial_bwd(const float* sa, const float* sb, const float* al, const float* be, const float* c,
                         const float* gs, float* gsa, float* gsb, int n, int dim, sfk_stream_t st);
int sfk_axpby(const float* x, const float* y, float* out, float a, float b, long n, sfk_stream_t st);

/* layout converters for the NCHW fp32 API surface */
int sfk_nchw_to_nhwc_bf16(const float* x, void* y, int n, int c, int h, int w, sfk_stream_t st);
int sfk_nhwc_bf16_to_nchw(const void* x, float* y, int n, int c, int h, int w, sfk_stream_t st);

/* ---------------------------------------------------------------------------------------------
 * Perturbation update (code/attack/interpolation.py:92-94 PGD step; adversarial_patch.py:131-138;
 * attack_main2.py:413-433 mask apply; attack_main2.py:606 Adam).  gpool is the gradient w.r.t. the
 * k x k box-pooled input; the full-resolution gradient is gscale*gpool[h/k][w/k].
 *   mode 0 linf : x = clamp(x0 + clamp(x + dir*alpha*sign(g) - x0, -eps, eps), lo, hi)
 *   mode 1 l2   : two-phase, see sfk_attack_update_l2
 *   mode 2 patch: patch += dir*lr*(sign?sign(g):g); x = clamp((1-m)*x0 + m*patch, lo[n], hi[n])
 *   mode 3 adam : torch.optim.Adam step on x (m, v state), no projection
 * stats[n] += sum |x_new - x0| (a cheap on-device progress metric reduced with warp shuffles). */
int sfk_attack_update_linf(float* x, const float* x0, const float* gpool, float alpha, float eps, float dir,
                           float lo, float hi, float* stats, int n, int size, int k, sfk_stream_t st);
/* PGD random start (interpolation.py:74-76): x = clamp(x0 + eps*U(-1,1), lo, hi), counter-based generator: element i of the
 * flattened tensor depends on (seed, i) only.  count % 4 == 0. */
int sfk_attack_random_start(float* x, const float* x0, float eps, float lo, float hi, unsigned long long seed, long count, sfk_stream_t st);
int sfk_attack_update_patch(float* x, const float* x0, float* patch, const float* mask, const float* gpool,
                            float lr, float dir, int use_sign, const float* lo, const float* hi, float gscale,
                            float* stats, int n, int size, int k, sfk_stream_t st);
/* gfull (optional): an additional full-resolution gradient term, g = gscale*gpool[h/k][w/k] + gfull_scale*gfull */
int sfk_attack_update_adam(float* x, const float* gpool, const float* gfull, float gfull_scale, float* m, float* v, float lr,
                           float b1, float b2, float eps, int t, float gscale, int n, int size, int k, sfk_stream_t st);
/* l2: norms[n] = sum g^2 (phase 0); x' = x + dir*alpha*g/|g|, dn[n] = sum (x'-x0)^2 (phase 1);
 *     x = clamp(x0 + (x'-x0)*min(1, eps/|d|), lo, hi) (phase 2) */
int sfk_attack_update_l2(float* x, const float* x0, const float* gpool, float* norms, float* dn, float alpha,
                         float eps, float dir, float lo, float hi, int phase, int n, int size, int k, sfk_stream_t st);
int sfk_minmax_per_sample(const float* x, float* lo, float* hi, int n, long per_sample, sfk_stream_t st);
/* Universal (shared) patch, the data-parallel form of patch.train (adversarial_patch.py:26-74): one patch / mask
 * (3 x size x size) for all n images.
 *   grad_reduce: gsum = mask * gscale * sum_n gpool[n][c][h/k][w/k]   (overwrites gsum; all-reduce it over ranks, then patch -= lr*gsum)
 *   apply:       x[n] = clamp((1-mask)*x0[n] + mask*patch, lo[n], hi[n])                    (adversarial_patch.py:137-138) */
int sfk_patch_grad_reduce(const float* gpool, const float* mask, float* gsum, float gscale, int n, int size, int k, sfk_stream_t st);
int sfk_patch_apply_shared(float* x, const float* x0, const float* patch, const float* mask, const float* lo, const float* hi, int n,
                           int size, sfk_stream_t st);

/* ---------------------------------------------------------------------------------------------
 * Outcome metric: SSIM of two RGB image batches (NCHW fp32), as cal_SSMI computes it (code/attack/interpolation.py:903-919:
 * skimage rgb2gray + structural_similarity with library defaults: 7x7 uniform window, sample covariance, K1 0.01, K2 0.03,
 * mean over the windows that lie inside the image).  out[n] = mean SSIM of image pair n.  data_range: 2 for [-1,1] floats. */
int sfk_ssim_gray7(const float* a, const float* b, float* out, int n, int h, int w, float data_range, sfk_stream_t st);

#ifdef __cplusplus
}
#endif
#endif
