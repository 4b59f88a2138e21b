"""ctypes binding of libsfattack.so (the C ABI declared in include/sfk.h).

No libtorch linkage: tensors cross the boundary as raw device pointers (``tensor.data_ptr()``) plus
sizes, launches go to torch's current CUDA stream.  There is NO fallback: if the library is missing
or a call returns non-zero, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libsfattack.so")

SFK_MAX_TAPS = 16
EP_DSCALE, EP_NOISE, EP_BIAS, EP_RELU, EP_LRELU, EP_XMASK, EP_GSDOT, EP_COLSCALE, EP_ACCUM, EP_LRELU_RAW = (
    1, 2, 4, 8, 16, 32, 64, 128, 256, 512)


class SfkTap(C.Structure):
    _fields_ = [("dy", C.c_int32), ("dx", C.c_int32), ("plane", C.c_int32), ("acc", C.c_int32), ("brow", C.c_int32)]


class SfkIgemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("n_img", C.c_int32), ("a_h", C.c_int32), ("a_w", C.c_int32), ("a_c", C.c_int32), ("a_planes", C.c_int32),
        ("b", C.c_void_p),
        ("b_samples", C.c_int32), ("b_rows", C.c_int32),
        ("out", C.c_void_p),
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("out_c", C.c_int32),
        ("num_acc", C.c_int32), ("block_n", C.c_int32),
        ("num_taps", C.c_int32),
        ("taps", SfkTap * SFK_MAX_TAPS),
        ("flags", C.c_int32),
        ("dscale", C.c_void_p), ("bias", C.c_void_p), ("noise", C.c_void_p),
        ("noise_w", C.c_float),
        ("xin", C.c_void_p), ("colscale", C.c_void_p), ("gs", C.c_void_p),
        ("vec_stride", C.c_int32),
        ("err", C.c_void_p),
        ("stages", C.c_int32),
        ("out_d2s", C.c_int32), ("a_s2d", C.c_int32),
        ("ws", C.c_void_p), ("ws_bytes", C.c_size_t),
    ]


# every symbol include/sfk.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "sfk_version", "sfk_set_activation_dtype", "sfk_get_activation_dtype", "sfk_last_error_string", "sfk_igemm", "sfk_igemm_prepare", "sfk_igemm_run", "sfk_igemm_destroy", "sfk_igemm_plan_info", "sfk_set_conv_math", "sfk_get_conv_math",
    "sfk_igemm_workspace_bytes", "sfk_igemm_v1", "sfk_role_cycles", "sfk_igemm_ref", "sfk_conv_c3_fwd", "sfk_conv_c3_bwd", "sfk_c3_pack", "sfk_c3_unpack",
    "sfk_avgpool_affine_fwd", "sfk_maxpool2_fwd", "sfk_maxpool2_bwd", "sfk_gap_fwd", "sfk_gap_bwd", "sfk_mse_tap",
    "sfk_mse_f32", "sfk_image_loss_grad", "sfk_style_affine_fwd", "sfk_style_affine_bwd", "sfk_demod_fwd", "sfk_demod_bwd",
    "sfk_modulate_weights", "sfk_demod_fwd_batched", "sfk_modulate_weights_batched", "sfk_demod_bwd_batched", "sfk_blur_act_fwd", "sfk_blur_act_bwd", "sfk_act_bwd", "sfk_torgb_fwd", "sfk_torgb_bwd", "sfk_act_torgb_bwd",
    "sfk_rgb_down", "sfk_linear_fwd", "sfk_linear_bwd", "sfk_fuse_spatial_fwd", "sfk_fuse_spatial_bwd", "sfk_axpby",
    "sfk_nchw_to_nhwc_bf16", "sfk_nhwc_bf16_to_nchw", "sfk_attack_update_linf", "sfk_attack_random_start", "sfk_attack_update_patch",
    "sfk_attack_update_adam", "sfk_attack_update_l2", "sfk_minmax_per_sample", "sfk_patch_grad_reduce", "sfk_patch_apply_shared",
    "sfk_ssim_gray7", "sfk_conv3x3_wgrad", "sfk_bias_grad", "sfk_modconv_wgrad_finish", "sfk_conv_c3_wgrad",
]

_lib = None


class SfkError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the in-tree library; raise loudly if it has not been built (no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SfkError(f"{LIB_PATH} not found: run `python __graft_entry__.py build` (the CUDA library is mandatory)")
        _lib = C.CDLL(LIB_PATH)
        _lib.sfk_last_error_string.restype = C.c_char_p
        _lib.sfk_version.restype = C.c_int
        _lib.sfk_igemm_workspace_bytes.restype = C.c_size_t
        _lib.sfk_igemm_run.argtypes = [C.c_void_p, C.c_void_p]
        _lib.sfk_igemm_destroy.argtypes = [C.c_void_p]
        _lib.sfk_igemm_plan_info.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def set_activation_dtype(dtype: torch.dtype):
    """torch.bfloat16 (default) or torch.float32 (parity mode).  Engines must be built AFTER switching."""
    assert dtype in (torch.bfloat16, torch.float32)
    load().sfk_set_activation_dtype(1 if dtype == torch.float32 else 0)


def activation_dtype() -> torch.dtype:
    return torch.float32 if load().sfk_get_activation_dtype() else torch.bfloat16


CONV_MATH = {"auto": 0, "tf32": 1, "tf32x3": 2, "cuda_cores": 3}


def set_conv_math(mode: str):
    """arithmetic of the tensor-core conv under fp32 storage (sfk.h): "tf32x3" (default: split tf32, fp32-class products),
    "tf32" (plain kind::tf32), "cuda_cores" (sfk_igemm_ref).  Plans are bound to the mode they were prepared under."""
    _chk0(load().sfk_set_conv_math(CONV_MATH[mode]), "set_conv_math")


def conv_math() -> str:
    m = load().sfk_get_conv_math()
    return {0: "tf32x3", 1: "tf32", 2: "tf32x3", 3: "cuda_cores"}[m]


def mode_key() -> tuple:
    """(storage dtype, conv math): everything process-global that a prepared plan or an engine's buffers depend on"""
    return (load().sfk_get_activation_dtype(), load().sfk_get_conv_math())


# scratch of the split-tf32 conv (one per device, shared by all launches: they are ordered on one stream).  Grown when a
# descriptor is created, never while plans that captured the old pointer are alive (the pointer is part of the plan key).
_WS = {}


def _workspace(dev: torch.device, need: int) -> torch.Tensor:
    t = _WS.get(dev)
    if t is None or t.numel() < need:
        t = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)
        _WS[dev] = t
    return t


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "sfk: tensors must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


LAUNCHES = 0          # kernels of libsfattack launched since last reset (every entry point launches exactly one)
_PROFILE = None       # when a list: igemm() appends (start_event, end_event, flops)


def _chk(rc: int, what: str):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        raise SfkError(f"{what} failed rc={rc}: {load().sfk_last_error_string().decode()}")


def _chk0(rc: int, what: str):
    if rc != 0:
        raise SfkError(f"{what} failed rc={rc}: {load().sfk_last_error_string().decode()}")


def _f(v) -> C.c_float:
    return C.c_float(float(v))


# ------------------------------------------------------------------------------------------------
def make_igemm_desc(a, n_img, a_h, a_w, a_c, a_planes, b, b_samples, b_rows, out, out_h, out_w, out_c, num_acc, block_n,
                    taps: Sequence[tuple], flags=0, dscale=None, bias=None, noise=None, noise_w=0.0, xin=None,
                    colscale=None, gs=None, err=None, stages=0, vec_stride=0, vec_off=0, out_d2s=0, a_s2d=0) -> SfkIgemmDesc:
    d = SfkIgemmDesc()
    block_n = fit_block_n(block_n, n_img, out_h, out_w, out_c)
    d.a, d.n_img, d.a_h, d.a_w, d.a_c, d.a_planes = _p(a), n_img, a_h, a_w, a_c, a_planes
    d.b, d.b_samples, d.b_rows = _p(b), b_samples, b_rows
    d.out, d.out_h, d.out_w, d.out_c = _p(out), out_h, out_w, out_c
    d.num_acc, d.block_n, d.num_taps = num_acc, block_n, len(taps)
    assert len(taps) <= SFK_MAX_TAPS
    for i, (dy, dx, plane, acc, brow) in enumerate(taps):
        d.taps[i] = SfkTap(dy, dx, plane, acc, brow)
    d.flags = flags
    d.dscale, d.bias, d.noise, d.noise_w = _p(dscale), _p(bias), _p(noise), float(noise_w)
    d.xin, d.err, d.stages, d.vec_stride = _p(xin), _p(err), stages, vec_stride
    d.out_d2s, d.a_s2d = out_d2s, a_s2d
    d.colscale = _sub(colscale, vec_off) if colscale is not None else C.c_void_p(0)
    d.gs = _sub(gs, vec_off) if gs is not None else C.c_void_p(0)
    need = int(load().sfk_igemm_workspace_bytes(C.byref(d)))
    if need:
        _workspace(a.device, need)
    d._dev = a.device
    d._need = need
    d._plan = None
    # keep the tensors alive as long as the descriptor
    d._keep = (a, b, out, dscale, bias, noise, xin, colscale, gs, err)
    return d


class _Plan:
    """owner of one sfk_igemm_plan (freed with the descriptor that cached it)"""

    def __init__(self, desc: "SfkIgemmDesc"):
        ws = _workspace(desc._dev, desc._need) if desc._need else None
        desc.ws, desc.ws_bytes = (ws.data_ptr(), ws.numel()) if ws is not None else (0, 0)
        h = C.c_void_p()
        _chk0(load().sfk_igemm_prepare(C.byref(desc), C.byref(h)), "sfk_igemm_prepare")
        self.h, self.key, self.ws = h, (mode_key(), desc.ws), ws

    def __del__(self):
        try:
            if self.h and _lib is not None:
                _lib.sfk_igemm_destroy(self.h)
        except Exception:
            pass


def igemm_flops(d: SfkIgemmDesc) -> float:
    """algorithmic FLOPs of one launch: every tap is an (out_c x a_c) MAC block per output position"""
    f = 2.0 * d.n_img * d.out_h * d.out_w * d.num_taps * d.out_c * d.a_c
    # fused upsample conv: the four phase convs execute 4x the MACs of the transposed conv they replace; count the latter
    return f / 4 if (d.out_d2s or d.a_s2d) else f


def _igemm_launch(desc: SfkIgemmDesc, ref: bool, v1: bool, oneshot: bool):
    L = load()
    if ref or v1:
        _chk((L.sfk_igemm_ref if ref else L.sfk_igemm_v1)(C.byref(desc), _stream()), "sfk_igemm_ref" if ref else "sfk_igemm_v1")
        return
    if oneshot:      # the plan-and-forget entry point (kept for callers that build descriptors on the fly)
        if desc._need:
            ws = _workspace(desc._dev, desc._need)
            desc.ws, desc.ws_bytes = ws.data_ptr(), ws.numel()
        _chk(L.sfk_igemm(C.byref(desc), _stream()), "sfk_igemm")
        return
    # prepared plan, cached on the descriptor; re-planned if the process-global modes or the workspace changed underneath it
    p = desc._plan
    if p is None or p.key[0] != mode_key() or (desc._need and _WS.get(desc._dev) is not p.ws):
        p = desc._plan = _Plan(desc)
    _chk(L.sfk_igemm_run(p.h, _stream()), "sfk_igemm_run")


PLAN_INFO_KEYS = ("m2", "halo", "resident", "d2s", "s2d", "passes", "stages", "block_n", "variant", "flags", "grid_x", "grid_y", "smem",
                  "f32", "cuda_cores", "acc_stages")


def plan_info(desc: SfkIgemmDesc) -> dict:
    """what the planner decided for this descriptor under the current modes (prepares the plan if needed)"""
    p = desc._plan
    if p is None or p.key[0] != mode_key() or (desc._need and _WS.get(desc._dev) is not p.ws):
        p = desc._plan = _Plan(desc)
    buf = (C.c_int32 * 16)()
    _chk0(load().sfk_igemm_plan_info(p.h, buf), "sfk_igemm_plan_info")
    d = dict(zip(PLAN_INFO_KEYS, list(buf)))
    d["ksplit"], d["acc_stages"] = d["acc_stages"] % 16, d["acc_stages"] // 16
    d.update(shape=(desc.n_img, desc.out_h, desc.out_w, desc.a_c, desc.out_c), taps=desc.num_taps, num_acc=desc.num_acc)
    return d


def igemm(desc: SfkIgemmDesc, ref: bool = False, v1: bool = False, oneshot: bool = False):
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _igemm_launch(desc, ref, v1, oneshot)
        e1.record()
        _PROFILE.append((e0, e1, igemm_flops(desc), (desc.n_img, desc.out_h, desc.out_w, desc.a_c, desc.out_c, desc.num_taps, desc.num_acc)))
        return
    _igemm_launch(desc, ref, v1, oneshot)


EP_PROFILE = 1 << 16


def role_cycles(reset=True):
    buf = (C.c_ulonglong * 8)()
    _chk(load().sfk_role_cycles(buf, int(reset)), "role_cycles")
    return list(buf)


def profile_igemm(step_fn) -> dict:
    """Run step_fn once with a CUDA-event pair around every tensor-core conv launch (on the launching stream)."""
    global _PROFILE
    _PROFILE = []
    try:
        step_fn()
        torch.cuda.synchronize()
        recs = [(a.elapsed_time(b), f, shp) for a, b, f, shp in _PROFILE]
    finally:
        _PROFILE = None
    return dict(ms=sum(r[0] for r in recs), flops=sum(r[1] for r in recs), launches=len(recs), per_launch=recs)


def conv3x3_taps(cout: int):
    """forward 3x3, pad 1: tap (ky,kx) reads x[h+ky-1][w+kx-1]; weight rows (ky*3+kx)*cout."""
    return [(ky - 1, kx - 1, 0, 0, (ky * 3 + kx) * cout) for ky in range(3) for kx in range(3)]


def conv3x3_dgrad_taps(cin: int):
    """data gradient of the same conv: gx[p] = sum_k gz[p-(k-1)] W[k]^T; weight rows are Wt[k][cin][cout]."""
    return [(1 - ky, 1 - kx, 0, 0, (ky * 3 + kx) * cin) for ky in range(3) for kx in range(3)]


def tconv_taps(cout: int):
    """stride-2 transposed 3x3 conv as 4 phase accumulators (tests/test_kernel_math.py): phase a=(ky==1), b=(kx==1);
    reads x[m-ky//2][n-kx//2]."""
    return [(-(ky // 2), -(kx // 2), 0, (1 if ky == 1 else 0) * 2 + (1 if kx == 1 else 0), (ky * 3 + kx) * cout)
            for ky in range(3) for kx in range(3)]


def tconv_dgrad_taps(cin: int):
    """transpose of the above: gx~[i] = sum_k gT[2i+k] W[k]^T, gT read from phase plane (ky%2, kx%2) at shift k//2."""
    return [(ky // 2, kx // 2, (ky % 2) * 2 + (kx % 2), 0, (ky * 3 + kx) * cin) for ky in range(3) for kx in range(3)]


def fit_block_n(block_n: int, n_img: int, out_h: int, out_w: int, out_c: int) -> int:
    """Narrower accumulators for launches that would leave most SMs idle.  A CTA owns one (image, N-block) and walks K serially,
    so the 4x4 .. 16x16 layers of the generator (K = 4608, one or two 128-row tiles per image) are bound by the length of that
    serial loop, not by bandwidth: 8 images x 4 N-blocks of 128 columns = 32 CTAs on 148 SMs, 72 k-steps of ~770 cycles each
    (round 1: 25-50 us per launch).  Halving block_n doubles the CTAs and shortens every k-step (the MMA's shared-memory reads
    are (128 + N) x 32 bytes).  SFK_LOWRES_BN=0 keeps the requested width."""
    if os.environ.get("SFK_LOWRES_BN", "1") == "0":
        return block_n
    tw = 16 if out_w > 8 else (8 if out_w > 4 else 4)
    tiles = -(-out_h // (128 // tw)) * -(-out_w // tw)
    while block_n > 32 and n_img * (out_c // block_n) * tiles < 100:
        block_n //= 2
    return block_n


def pick_block_n(cout: int, num_acc: int = 1) -> int:
    """columns per accumulator: 128 for single-accumulator launches; 64 for the 4-phase transposed conv so that the four
    accumulators (256 columns) can be double-buffered in TMEM (65^2 512->256: 200 -> 160 us)"""
    bn = min(cout, 128 if num_acc == 1 else 64)
    while num_acc * bn > 512:
        bn //= 2
    return bn


# ------------------------------------------------------------------------------------------------
def conv_c3_fwd(x, w, bias, out, relu=True):
    n, _, h, wd = x.shape
    _chk(load().sfk_conv_c3_fwd(_p(x), _p(w), _p(bias), _p(out), n, h, wd, w.shape[0], int(relu), _stream()), "conv_c3_fwd")


def conv_c3_bwd(g, w, gx):
    n, _, h, wd = gx.shape
    _chk(load().sfk_conv_c3_bwd(_p(g), _p(w), _p(gx), n, h, wd, w.shape[0], _stream()), "conv_c3_bwd")


def c3_pack(x, xp):
    """x (n,3,h,w) fp32 -> xp (n,h,w,16) bf16 = [hi(3) | lo(3) | hi(3) | 0 x 7], the tensor-core operand of the first conv"""
    n, _, h, wd = x.shape
    _chk(load().sfk_c3_pack(_p(x), _p(xp), n, h, wd, _stream()), "c3_pack")


def c3_unpack(gp, gx):
    """gp (n,h,w,16) bf16 (data gradient of the packed operand) -> gx (n,3,h,w) fp32 = channels [0:3] + [3:6]"""
    n, _, h, wd = gx.shape
    _chk(load().sfk_c3_unpack(_p(gp), _p(gx), n, h, wd, _stream()), "c3_unpack")


def c3_pack_weights(W: torch.Tensor):
    """W (cout,3,3,3) fp32 -> (forward weights [9*cout][16], data-gradient weights [9*16][cout]) in bf16 for the packed operand:
    forward K-channels [W.hi | W.hi | W.lo | 0]; gradient rows [W.hi^T (3) | W.lo^T (3) | 0 x 10] per tap."""
    cout = W.shape[0]
    W = W.to(torch.float32)
    hi = W.bfloat16().float()
    lo = (W - hi).bfloat16().float()
    wf = torch.zeros(3, 3, cout, 16, dtype=torch.float32, device=W.device)
    hik, lok = hi.permute(2, 3, 0, 1), lo.permute(2, 3, 0, 1)          # (ky,kx,cout,3)
    wf[..., 0:3], wf[..., 3:6], wf[..., 6:9] = hik, hik, lok
    wb = torch.zeros(3, 3, 16, cout, dtype=torch.float32, device=W.device)
    wb[:, :, 0:3], wb[:, :, 3:6] = hi.permute(2, 3, 1, 0), lo.permute(2, 3, 1, 0)
    return wf.reshape(9 * cout, 16).bfloat16().contiguous(), wb.reshape(9 * 16, cout).bfloat16().contiguous()


def avgpool_affine_fwd(x, y, k, a=1.0, b=0.0):
    planes = x.shape[0] * x.shape[1]
    _chk(load().sfk_avgpool_affine_fwd(_p(x), _p(y), planes, x.shape[2], x.shape[3], k, _f(a), _f(b), _stream()), "avgpool")


def maxpool2_fwd(x, y):
    n, h, w, c = x.shape
    _chk(load().sfk_maxpool2_fwd(_p(x), _p(y), n, h, w, c, _stream()), "maxpool2_fwd")


def maxpool2_bwd(x, gy, gx, tap_ref=None, tap_coef=0.0, relu_mask=True):
    n, h, w, c = x.shape
    _chk(load().sfk_maxpool2_bwd(_p(x), C.c_void_p(0), _p(gy), _p(gx), _p(tap_ref), _f(tap_coef), int(relu_mask), n, h, w, c,
                                 _stream()), "maxpool2_bwd")


def gap_fwd(x, y):
    n, h, w, c = x.shape
    _chk(load().sfk_gap_fwd(_p(x), _p(y), n, h * w, c, _stream()), "gap_fwd")


def gap_bwd(x, gy, gx):
    n, h, w, c = x.shape
    _chk(load().sfk_gap_bwd(_p(x), _p(gy), _p(gx), n, h * w, c, _stream()), "gap_bwd")


def mse_tap(f, ref, g, loss, coef_loss, coef_grad, accumulate=False, relu_mask=False):
    n = f.shape[0]
    per = f.numel() // n
    _chk(load().sfk_mse_tap(_p(f), _p(ref), _p(g), _p(loss), _f(coef_loss), _f(coef_grad), int(accumulate), int(relu_mask), n,
                            C.c_long(per), _stream()), "mse_tap")


def mse_f32(a, b, g, loss, coef_loss, coef_grad, accumulate=False):
    n = a.shape[0]
    _chk(load().sfk_mse_f32(_p(a), _p(b), _p(g), _p(loss), _f(coef_loss), _f(coef_grad), int(accumulate), n, C.c_long(a.numel() // n), _stream()),
         "mse_f32")


def image_loss_grad(img, ref, gpool, g, loss, coef_loss, coef_grad, k):
    n, _, s, _ = img.shape
    _chk(load().sfk_image_loss_grad(_p(img), _p(ref), _p(gpool), _p(g), _p(loss), _f(coef_loss), _f(coef_grad), n, s, k, _stream()),
         "image_loss_grad")


def style_affine_fwd(w, A, bias, row_widx, s, scale):
    n, L, D = w.shape
    _chk(load().sfk_style_affine_fwd(_p(w), _p(A), _p(bias), _p(row_widx), _p(s), n, L, D, A.shape[0], _f(scale), _stream()),
         "style_affine_fwd")


def style_affine_bwd(gs, A, layer_row_start, layer_widx, gw, scale):
    n, L, D = gw.shape
    _chk(load().sfk_style_affine_bwd(_p(gs), _p(A), _p(layer_row_start), _p(layer_widx), layer_widx.numel(), _p(gw), n, L, D,
                                     A.shape[0], _f(scale), _stream()), "style_affine_bwd")


def _sub(t: torch.Tensor, off: int) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() + off * t.element_size())


def demod_fwd(s, s_off, Q, d):
    n, sd = s.shape
    cout, cin = Q.shape
    _chk(load().sfk_demod_fwd(_sub(s, s_off), sd, _p(Q), _p(d), n, cin, cout, _stream()), "demod_fwd")


def demod_bwd(s, s_off, Q, d, gdacc, gs):
    n, sd = s.shape
    cout, cin = Q.shape
    _chk(load().sfk_demod_bwd(_sub(s, s_off), sd, _p(Q), _p(d), _p(gdacc), _sub(gs, s_off), gs.shape[1], n, cin, cout, _stream()),
         "demod_bwd")


def modulate_weights(wbase, s, s_off, wmod, d=None):
    n, sd = s.shape
    taps, cout, cin = wbase.shape
    _chk(load().sfk_modulate_weights(_p(wbase), _sub(s, s_off), sd, _p(wmod), n, taps, cout, cin, _p(d),
                                     d.shape[1] if d is not None else 0, _stream()), "modulate_weights")


STYLE_TAB_COLS = 10   # s_off, cin, cout, q_off, d_off, rows, wb_off, wm_off, d_cols, fold  (int64, see sfk.h)


def demod_fwd_batched(s, q_cat, d_cat, tab, max_cout):
    n, sd = s.shape
    _chk(load().sfk_demod_fwd_batched(_p(s), sd, _p(q_cat), _p(d_cat), _p(tab), tab.shape[0], n, max_cout, _stream()), "demod_fwd_batched")


def modulate_weights_batched(wbase_cat, s, wmod_cat, d_cat, tab):
    n, sd = s.shape
    _chk(load().sfk_modulate_weights_batched(_p(wbase_cat), _p(s), sd, _p(wmod_cat), _p(d_cat), _p(tab), tab.shape[0], n, _stream()),
         "modulate_weights_batched")


def demod_bwd_batched(s, q_cat, d_cat, gd_cat, gs, tab, max_cin):
    n, sd = s.shape
    _chk(load().sfk_demod_bwd_batched(_p(s), sd, _p(q_cat), _p(d_cat), _p(gd_cat), _p(gs), gs.shape[1], _p(tab), tab.shape[0], n, max_cin,
                                      _stream()), "demod_bwd_batched")


def conv3x3_wgrad(x, gz, dw=None, per_sample=False, ref=False, err=None):
    """dw (S,9,cout,cin) fp32 += sum gz (x) x shifted by the tap; x (n,h,w,cin), gz (n,h,w,cout) NHWC activations (sfk.h).
    Returns dw (allocated and zeroed when not given)."""
    n, h, w, cin = x.shape
    cout = gz.shape[3]
    if dw is None:
        dw = torch.zeros((n if per_sample else 1), 9, cout, cin, device=x.device, dtype=torch.float32)
    _chk(load().sfk_conv3x3_wgrad(_p(x), _p(gz), _p(dw), n, h, w, cin, cout, int(per_sample), int(ref),
                                  _p(err) if err is not None else C.c_void_p(0), _stream()), "conv3x3_wgrad")
    return dw


def conv_c3_wgrad(x, gz, dw=None):
    """x (n,3,h,w) fp32 NCHW, gz (n,h,w,cout) -> dw (cout,3,3,3) fp32 (torch layout), accumulated"""
    n, _, h, w = x.shape
    cout = gz.shape[3]
    if dw is None:
        dw = torch.zeros(cout, 3, 3, 3, device=x.device, dtype=torch.float32)
    _chk(load().sfk_conv_c3_wgrad(_p(x), _p(gz), _p(dw), n, h, w, cout, _stream()), "conv_c3_wgrad")
    return dw


def bias_grad(gz, db=None):
    n, h, w, c = gz.shape
    if db is None:
        db = torch.zeros(c, device=gz.device, dtype=torch.float32)
    _chk(load().sfk_bias_grad(_p(gz), _p(db), n, h * w, c, _stream()), "bias_grad")
    return db


def modconv_wgrad_finish(G, wbase, s, s_off, d, gdacc, demodulate=True):
    """G (n,9,cout,cin) from conv3x3_wgrad(per_sample=True) -> gradient of the shared base weight (9,cout,cin)."""
    n, _, cout, cin = G.shape
    dwb = torch.empty(9, cout, cin, device=G.device, dtype=torch.float32)
    _chk(load().sfk_modconv_wgrad_finish(_p(G), _p(wbase), _sub(s, s_off), s.shape[1], _p(d) if d is not None else C.c_void_p(0),
                                         _p(gdacc) if gdacc is not None else C.c_void_p(0), _p(dwb), n, cout, cin, int(demodulate),
                                         _stream()), "modconv_wgrad_finish")
    return dwb


def blur_act_fwd(T, out, d, noise, noise_w, bias):
    n, ho, wo, c = out.shape
    _chk(load().sfk_blur_act_fwd(_p(T), _p(out), _p(d), _p(noise), _f(noise_w), _p(bias), n, ho // 2, wo // 2, c, _stream()),
         "blur_act_fwd")


def blur_act_bwd(out, gout, gT, d, noise, noise_w, bias, gdacc, s_in=None, gs_in=None, in_off=0):
    """s_in / gs_in (B, s_dim) + in_off: finish a flags-0 data gradient of the conv that consumes `out` (as act_bwd does)."""
    n, ho, wo, c = out.shape
    _chk(load().sfk_blur_act_bwd(_p(out), _p(gout), _p(gT), _p(d), _p(noise), _f(noise_w), _p(bias), _p(gdacc),
                                 _sub(s_in, in_off) if s_in is not None else C.c_void_p(0),
                                 _sub(gs_in, in_off) if gs_in is not None else C.c_void_p(0),
                                 s_in.shape[1] if s_in is not None else 0, n, ho // 2, wo // 2, c, _stream()), "blur_act_bwd")


def act_bwd(out, gout, gz, d, noise, noise_w, bias, gdacc, s_in=None, gs_in=None, in_off=0):
    """s_in / gs_in (B, s_dim) + in_off: finish a flags-0 data gradient of the conv that consumes `out` (see sfk.h)."""
    n, h, w, c = out.shape
    _chk(load().sfk_act_bwd(_p(out), _p(gout), _p(gz), _p(d), _p(noise), _f(noise_w), _p(bias), _p(gdacc),
                            _sub(s_in, in_off) if s_in is not None else C.c_void_p(0),
                            _sub(gs_in, in_off) if gs_in is not None else C.c_void_p(0),
                            s_in.shape[1] if s_in is not None else 0, n, h, w, c, _stream()), "act_bwd")


def torgb_fwd(x, wrgb, s, s_off, bias, skip, rgb):
    n, h, w, c = x.shape
    _chk(load().sfk_torgb_fwd(_p(x), _p(wrgb), _sub(s, s_off), s.shape[1], _p(bias), _p(skip), _p(rgb), n, h, w, c, _stream()),
         "torgb_fwd")


def torgb_bwd(x, wrgb, s, s_off, grgb, gx, gs):
    n, h, w, c = x.shape
    _chk(load().sfk_torgb_bwd(_p(x), _p(wrgb), _sub(s, s_off), s.shape[1], _p(grgb), _p(gx), _sub(gs, s_off), gs.shape[1], n, h, w,
                              c, _stream()), "torgb_bwd")


def act_torgb_bwd(out, gin, gz, d, noise, noise_w, bias, gdacc, wrgb, s, s_off, grgb, gs, in_off=None):
    """in_off: s_off of the conv whose flags-0 data gradient wrote `gin` (its modulation and style gradient are finished here)."""
    n, h, w, c = out.shape
    fin = in_off is not None and gin is not None
    _chk(load().sfk_act_torgb_bwd(_p(out), _p(gin), _p(gz), _p(d), _p(noise), _f(noise_w), _p(bias), _p(gdacc), _p(wrgb), _sub(s, s_off),
                                  s.shape[1], _p(grgb), _sub(gs, s_off), gs.shape[1], _sub(s, in_off) if fin else C.c_void_p(0),
                                  _sub(gs, in_off) if fin else C.c_void_p(0), n, h, w, c, _stream()), "act_torgb_bwd")


def rgb_down(g, gskip):
    n, c, h, w = g.shape
    _chk(load().sfk_rgb_down(_p(g), _p(gskip), n * c, h, w, _stream()), "rgb_down")


def linear_fwd(x, W, bias, y):
    _chk(load().sfk_linear_fwd(_p(x), _p(W), _p(bias), _p(y), x.shape[0], W.shape[1], W.shape[0], _stream()), "linear_fwd")


def linear_bwd(gy, W, gx):
    _chk(load().sfk_linear_bwd(_p(gy), _p(W), _p(gx), gy.shape[0], W.shape[1], W.shape[0], _stream()), "linear_bwd")


def fuse_spatial_fwd(sa, sb, al, be, c, s):
    _chk(load().sfk_fuse_spatial_fwd(_p(sa), _p(sb), _p(al), _p(be), _p(c), _p(s), sa.shape[0], sa.shape[1], _stream()), "fuse_fwd")


def fuse_spatial_bwd(sa, sb, al, be, c, gs, gsa, gsb):
    _chk(load().sfk_fuse_spatial_bwd(_p(sa), _p(sb), _p(al), _p(be), _p(c), _p(gs), _p(gsa), _p(gsb), sa.shape[0], sa.shape[1],
                                     _stream()), "fuse_bwd")


def axpby(x, y, out, a, b=0.0):
    _chk(load().sfk_axpby(_p(x), _p(y), _p(out), _f(a), _f(b), C.c_long(x.numel()), _stream()), "axpby")


def nchw_to_nhwc_bf16(x, y):
    n, c, h, w = x.shape
    _chk(load().sfk_nchw_to_nhwc_bf16(_p(x), _p(y), n, c, h, w, _stream()), "nchw_to_nhwc")


def nhwc_bf16_to_nchw(x, y):
    n, h, w, c = x.shape
    _chk(load().sfk_nhwc_bf16_to_nchw(_p(x), _p(y), n, c, h, w, _stream()), "nhwc_to_nchw")


def attack_update_linf(x, x0, gpool, alpha, eps, direction, lo, hi, stats, k):
    n, _, s, _ = x.shape
    _check_update_args(x, gpool, k, x0=x0)
    _chk(load().sfk_attack_update_linf(_p(x), _p(x0), _p(gpool), _f(alpha), _f(eps), _f(direction), _f(lo), _f(hi), _p(stats), n, s, k,
                                       _stream()), "attack_update_linf")


def attack_random_start(x, x0, eps, seed, lo=0.0, hi=1.0):
    """x <- clamp(x0 + eps*U(-1,1), lo, hi) on the device (interpolation.py:74-76); element i depends on (seed, i) only."""
    assert x.shape == x0.shape and x.is_contiguous() and x0.is_contiguous() and x.dtype == torch.float32
    _chk(load().sfk_attack_random_start(_p(x), _p(x0), _f(eps), _f(lo), _f(hi), C.c_ulonglong(int(seed) & (2 ** 64 - 1)), C.c_long(x.numel()),
                                        _stream()), "attack_random_start")


def _check_update_args(x, gpool, k, **same_shape):
    """the update kernels index every per-image operand with the image index: a broadcastable-but-smaller tensor would be read
    out of bounds, so shapes are checked here (the C ABI only sees pointers)"""
    n, c, s, s2 = x.shape
    assert c == 3 and s == s2 and s % k == 0, f"sfk: image batch must be (n,3,S,S) with S % k == 0, got {tuple(x.shape)}, k={k}"
    assert tuple(gpool.shape) == (n, 3, s // k, s // k), f"sfk: pooled gradient {tuple(gpool.shape)} does not match {tuple(x.shape)} / {k}"
    for name, t in same_shape.items():
        assert t is None or tuple(t.shape) == tuple(x.shape), f"sfk: `{name}` {tuple(t.shape)} must have the shape of x {tuple(x.shape)}"


def attack_update_patch(x, x0, patch, mask, gpool, lr, direction, use_sign, lo, hi, gscale, stats, k):
    n, _, s, _ = x.shape
    _check_update_args(x, gpool, k, x0=x0, patch=patch, mask=mask)
    assert lo.numel() >= n and hi.numel() >= n
    _chk(load().sfk_attack_update_patch(_p(x), _p(x0), _p(patch), _p(mask), _p(gpool), _f(lr), _f(direction), int(use_sign), _p(lo),
                                        _p(hi), _f(gscale), _p(stats), n, s, k, _stream()), "attack_update_patch")


def attack_update_adam(x, gpool, m, v, lr, t, gscale, k, b1=0.9, b2=0.999, eps=1e-8, gfull=None, gfull_scale=1.0):
    n, _, s, _ = x.shape
    _check_update_args(x, gpool, k, m=m, v=v, gfull=gfull)
    _chk(load().sfk_attack_update_adam(_p(x), _p(gpool), _p(gfull), _f(gfull_scale), _p(m), _p(v), _f(lr), _f(b1), _f(b2), _f(eps), int(t), _f(gscale), n, s, k,
                                       _stream()), "attack_update_adam")


def attack_update_l2(x, x0, gpool, norms, dn, alpha, eps, direction, lo, hi, phase, k):
    n, _, s, _ = x.shape
    _check_update_args(x, gpool, k, x0=x0)
    assert norms.numel() >= n and dn.numel() >= n
    _chk(load().sfk_attack_update_l2(_p(x), _p(x0), _p(gpool), _p(norms), _p(dn), _f(alpha), _f(eps), _f(direction), _f(lo), _f(hi),
                                     int(phase), n, s, k, _stream()), "attack_update_l2")


def minmax_per_sample(x, lo, hi):
    n = x.shape[0]
    _chk(load().sfk_minmax_per_sample(_p(x), _p(lo), _p(hi), n, C.c_long(x.numel() // n), _stream()), "minmax")


def patch_grad_reduce(gpool, mask, gsum, gscale, k):
    """gsum (1,3,S,S) = mask * gscale * sum_n gpool[n][c][h/k][w/k]: gradient of the batch w.r.t. ONE shared patch"""
    n, _, sp, _ = gpool.shape
    s = sp * k
    assert tuple(mask.shape[-3:]) == (3, s, s) and tuple(gsum.shape[-3:]) == (3, s, s) and mask.numel() == gsum.numel() == 3 * s * s
    _chk(load().sfk_patch_grad_reduce(_p(gpool), _p(mask), _p(gsum), _f(gscale), n, s, k, _stream()), "patch_grad_reduce")


def patch_apply_shared(x, x0, patch, mask, lo, hi):
    n, _, s, _ = x.shape
    assert tuple(x0.shape) == tuple(x.shape) and patch.numel() == mask.numel() == 3 * s * s and lo.numel() >= n and hi.numel() >= n
    _chk(load().sfk_patch_apply_shared(_p(x), _p(x0), _p(patch), _p(mask), _p(lo), _p(hi), n, s, _stream()), "patch_apply_shared")


def ssim_gray7(a, b, out, data_range=2.0):
    n, c, h, w = a.shape
    assert c == 3 and tuple(b.shape) == tuple(a.shape) and out.numel() >= n
    _chk(load().sfk_ssim_gray7(_p(a), _p(b), _p(out), n, h, w, _f(data_range), _stream()), "ssim_gray7")
