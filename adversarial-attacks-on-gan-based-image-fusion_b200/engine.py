"""The attack hot path as explicit forward/backward kernel schedules over pre-allocated HBM buffers.

Nothing here computes: every arithmetic step is a launch into libsfattack.so (lib.py).  torch is used
for device memory and streams only.  Three schedules:

  ConvStack        3x3 conv + bias + ReLU / 2x2 max-pool chains: the reference's VGG feature extractor
                   (code/vgg.py:44-64) and the encoder stand-in.
  SynthesisEngine  StyleGAN2 config-f synthesis from StyleSpace vectors, forward and data/style backward
                   (the `decoder([w], input_is_latent=True, ...)` call, code/attack/attack_main2.py:619-621).
  AttackEngine     encoder -> pair fusion -> synthesis -> VGG/pixel loss -> backward -> perturbation update
                   for a batch of independent image pairs (the loops of attack_main2.py:614-653 and
                   attack/patch/adversarial_patch.py:111-158, re-posed on the fused output as north_star asks).

Layouts: activations NHWC bf16, images NCHW fp32, style vectors fp32.
"""
from __future__ import annotations

import functools
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib
from .params import EncSpec, GenSpec, VGG_CONVS, fused_up_base_weights

def ACT() -> torch.dtype:
    """activation storage type of the library right now: bf16 (product) or fp32 (parity mode)"""
    return lib.activation_dtype()


def _on_device(fn):
    """Launches go to torch's current stream of the CURRENT device; an engine built for another device switches to it first
    (the C ABI only sees pointers and a stream)."""
    @functools.wraps(fn)
    def wrap(self, *a, **k):
        dev = self.dev
        if dev.type == "cuda" and dev.index is not None and torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return fn(self, *a, **k)
        return fn(self, *a, **k)
    return wrap


def _empty(shape, dev, dtype=None):
    return torch.empty(shape, device=dev, dtype=ACT() if dtype is None else dtype)


def _zeros(shape, dev, dtype=torch.float32):
    return torch.zeros(shape, device=dev, dtype=dtype)


# =================================================================================================
@dataclass
class StackLayer:
    kind: str            # "c3" | "conv" | "pool"
    cin: int = 0
    cout: int = 0
    tap: int = -1        # index into the tap list, -1 = not a tap
    h: int = 0           # output spatial size (filled by ConvStack)
    w: int = 0


def vgg_layers(width_div: int = 1) -> List[StackLayer]:
    """conv1_1[tap0] conv1_2[tap1] pool1 conv2_1 conv2_2 pool2[tap2, named "conv3_2" by the reference]
    conv3_1 conv3_2 conv3_3 pool3(ceil) conv4_1 conv4_2[tap3]   (code/vgg.py:44-64)."""
    c = lambda v: v // width_div
    return [StackLayer("c3", 3, c(64), tap=0), StackLayer("conv", c(64), c(64), tap=1), StackLayer("pool"),
            StackLayer("conv", c(64), c(128)), StackLayer("conv", c(128), c(128)), StackLayer("pool", tap=2),
            StackLayer("conv", c(128), c(256)), StackLayer("conv", c(256), c(256)), StackLayer("conv", c(256), c(256)),
            StackLayer("pool"), StackLayer("conv", c(256), c(512)), StackLayer("conv", c(512), c(512), tap=3)]


def encoder_layers(spec: EncSpec) -> List[StackLayer]:
    L: List[StackLayer] = []
    cin = 3
    for i, co in enumerate(spec.widths):
        L.append(StackLayer("c3" if i == 0 else "conv", cin, co))
        if i < len(spec.widths) - 1:
            L.append(StackLayer("pool"))
        cin = co
    return L


class ConvStack:
    def __init__(self, layers: List[StackLayer], weights: Sequence[Tuple[torch.Tensor, torch.Tensor]], n: int, res: int,
                 device, err: torch.Tensor):
        self.layers, self.n, self.res, self.dev, self.err = layers, n, res, torch.device(device), err
        self.mode = lib.mode_key()      # storage dtype / conv math the buffers and plans were built for
        self.w_f32, self.bias, self.w_fwd, self.w_bwd = {}, {}, {}, {}
        # first conv (3 input channels) on the tensor cores: the image travels as a 16-channel hi/lo-split bf16 operand
        # (lib.c3_pack; 8 images at 256^2: 85 -> 30 us forward, 95 -> 25 us backward).  bf16 storage only; SFK_C3_TC=0 keeps the
        # CUDA-core kernels (which the fp32 parity mode always runs).
        self.c3_tc = ACT() == torch.bfloat16 and os.environ.get("SFK_C3_TC", "1") != "0" and layers[0].kind == "c3" and \
            layers[0].cout % 16 == 0
        h = w = res
        c = 3
        wi = 0
        self.out: List[torch.Tensor] = []
        self.g: List[torch.Tensor] = []
        for i, l in enumerate(layers):
            if l.kind == "pool":
                h, w = (h + 1) // 2, (w + 1) // 2
                l.cin = l.cout = c
            else:
                W, b = weights[wi]
                wi += 1
                assert W.shape[0] == l.cout and W.shape[1] == l.cin, (W.shape, l)
                self.bias[i] = b.to(device=device, dtype=torch.float32).contiguous()
                if l.kind == "c3":
                    self.w_f32[i] = W.to(device=device, dtype=torch.float32).contiguous()
                    if self.c3_tc:
                        self.w_fwd[i], self.w_bwd[i] = lib.c3_pack_weights(self.w_f32[i])
                else:
                    Wd = W.to(device=device, dtype=torch.float32)
                    self.w_fwd[i] = Wd.permute(2, 3, 0, 1).reshape(9 * l.cout, l.cin).to(ACT()).contiguous()   # [tap][cout][cin]
                    self.w_bwd[i] = Wd.permute(2, 3, 1, 0).reshape(9 * l.cin, l.cout).to(ACT()).contiguous()   # [tap][cin][cout]
                c = l.cout
            l.h, l.w = h, w
            self.out.append(_empty((n, h, w, c), device))
            self.g.append(_empty((n, h, w, c), device))
        self.g_in = _empty((n, 3, res, res), device, torch.float32)
        if self.c3_tc:
            self.xp = _empty((n, res, res, 16), device)       # packed image operand
            self.gp = _empty((n, res, res, 16), device)       # its data gradient
        self.taps = [i for i, l in enumerate(layers) if l.tap >= 0]
        self._fwd_desc, self._bwd_desc = {}, {}
        self._build_descs()

    def _build_descs(self):
        n = self.n
        for i, l in enumerate(self.layers):
            if l.kind == "c3" and self.c3_tc:
                assert i == 0
                self._fwd_desc[i] = lib.make_igemm_desc(
                    self.xp, n, l.h, l.w, 16, 1, self.w_fwd[i], 1, 9 * l.cout, self.out[i], l.h, l.w, l.cout, 1,
                    lib.pick_block_n(l.cout), lib.conv3x3_taps(l.cout), flags=lib.EP_BIAS | lib.EP_RELU, bias=self.bias[i], err=self.err)
                self._bwd_desc[i] = lib.make_igemm_desc(
                    self.g[i], n, l.h, l.w, l.cout, 1, self.w_bwd[i], 1, 9 * 16, self.gp, l.h, l.w, 16, 1, 16,
                    lib.conv3x3_dgrad_taps(16), err=self.err)
                continue
            if l.kind != "conv":
                continue
            prev = self.layers[i - 1]
            self._fwd_desc[i] = lib.make_igemm_desc(
                self.out[i - 1], n, prev.h, prev.w, l.cin, 1, self.w_fwd[i], 1, 9 * l.cout, self.out[i], l.h, l.w, l.cout, 1,
                lib.pick_block_n(l.cout), lib.conv3x3_taps(l.cout), flags=lib.EP_BIAS | lib.EP_RELU, bias=self.bias[i], err=self.err)
            flags = lib.EP_XMASK if prev.kind != "pool" else 0
            self._bwd_desc[i] = lib.make_igemm_desc(
                self.g[i], n, l.h, l.w, l.cout, 1, self.w_bwd[i], 1, 9 * l.cin, self.g[i - 1], prev.h, prev.w, l.cin, 1,
                lib.pick_block_n(l.cin), lib.conv3x3_dgrad_taps(l.cin), flags=flags,
                xin=self.out[i - 1] if flags else None, err=self.err)

    @_on_device
    def forward(self, x: torch.Tensor):
        """x: (n,3,res,res) fp32 NCHW."""
        for i, l in enumerate(self.layers):
            if l.kind == "c3" and self.c3_tc:
                lib.c3_pack(x, self.xp)
                lib.igemm(self._fwd_desc[i])
            elif l.kind == "c3":
                lib.conv_c3_fwd(x, self.w_f32[i], self.bias[i], self.out[i], relu=True)
            elif l.kind == "conv":
                lib.igemm(self._fwd_desc[i])
            else:
                lib.maxpool2_fwd(self.out[i - 1], self.out[i])
        return self.out[-1]

    def tap_outputs(self) -> List[torch.Tensor]:
        return [self.out[i] for i in self.taps]

    @_on_device
    def weight_grads(self, x: torch.Tensor) -> Dict[int, Tuple[torch.Tensor, torch.Tensor]]:
        """The second backward GEMM (the reference's autograd runs it because it never freezes parameters,
        attack_main2.py:301-304): {layer index: (dW (cout,cin,3,3), db (cout,))} in fp32, from the buffers a forward(x) + backward()
        left behind -- g[i] is dL/d(pre-activation) of conv i (the ReLU mask is applied by whoever wrote it), out[i-1] its input."""
        res = {}
        for i, l in enumerate(self.layers):
            if l.kind == "c3":
                dw = lib.conv_c3_wgrad(x, self.g[i])
            elif l.kind == "conv":
                dw = lib.conv3x3_wgrad(self.out[i - 1], self.g[i], err=self.err)[0]
                dw = dw.view(3, 3, l.cout, l.cin).permute(2, 3, 0, 1).contiguous()
            else:
                continue
            res[i] = (dw, lib.bias_grad(self.g[i]))
        return res

    def _tap_coefs(self, i: int, coef: float):
        per = self.out[i].numel() // self.n
        return coef / per, 2.0 * coef / per

    @_on_device
    def backward(self, tap_refs: Optional[Sequence[torch.Tensor]] = None, tap_coef: float = 0.0,
                 loss: Optional[torch.Tensor] = None, top_grad_ready: bool = False) -> torch.Tensor:
        """Back-propagate sum_taps coef*MSE(tap, ref) (and/or a gradient already stored, masked, in g[-1])
        down to the stack input.  Returns g_in (n,3,res,res) fp32 (owned by the stack)."""
        L = self.layers
        last = len(L) - 1

        def tap_at(i):
            return tap_refs is not None and L[i].tap >= 0

        if not top_grad_ready:
            assert tap_at(last)
            cl, cg = self._tap_coefs(last, tap_coef)
            lib.mse_tap(self.out[last], tap_refs[L[last].tap], self.g[last], loss, cl, cg, accumulate=False,
                        relu_mask=L[last].kind != "pool")
        for i in range(last, -1, -1):
            l = L[i]
            if l.kind == "conv":
                lib.igemm(self._bwd_desc[i])
                if tap_at(i - 1):
                    cl, cg = self._tap_coefs(i - 1, tap_coef)
                    lib.mse_tap(self.out[i - 1], tap_refs[L[i - 1].tap], self.g[i - 1], loss, cl, cg, accumulate=True,
                                relu_mask=L[i - 1].kind != "pool")
            elif l.kind == "pool":
                if tap_at(i - 1):
                    cl, cg = self._tap_coefs(i - 1, tap_coef)
                    lib.mse_tap(self.out[i - 1], tap_refs[L[i - 1].tap], None, loss, cl, 0.0)   # loss value only
                    lib.maxpool2_bwd(self.out[i - 1], self.g[i], self.g[i - 1], tap_ref=tap_refs[L[i - 1].tap], tap_coef=cg,
                                     relu_mask=True)
                else:
                    lib.maxpool2_bwd(self.out[i - 1], self.g[i], self.g[i - 1], relu_mask=True)
            elif self.c3_tc:
                lib.igemm(self._bwd_desc[i])
                lib.c3_unpack(self.gp, self.g_in)
            else:  # c3
                lib.conv_c3_bwd(self.g[i], self.w_f32[i], self.g_in)
        return self.g_in

    @property
    def launches_fwd(self):
        return len(self.layers)


# =================================================================================================
def _fused_up_min_res() -> int:
    """Up-layers at or above this output resolution run as ONE fused upsample-conv launch (2.25x the tensor FLOPs, but no
    phase-plane round trip and no separate blur kernels); below it the 4-accumulator transposed conv + the streaming blur kernels
    are cheaper.  Measured on the 1024 model, 8 pairs (ms/step): round 1 (tiled blur, 1.2 TB/s) none 14.0, >=1024 13.28,
    >=512 12.79, >=256 12.63; round 2 (streaming blur, 3-4 TB/s) >=256 8.45, >=512 8.28, >=1024 8.20, none 8.23.
    SFK_FUSED_UP_RES overrides (0 = all layers, large = none)."""
    return int(os.environ.get("SFK_FUSED_UP_RES", "1024"))


class SynthesisEngine:
    """StyleGAN2 synthesis from a concatenated StyleSpace vector s (B, s_dim)."""

    def __init__(self, spec: GenSpec, P: Dict[str, torch.Tensor], batch: int, device, err: torch.Tensor):
        self.spec, self.B, self.dev, self.err = spec, batch, torch.device(device), err
        self.mode = lib.mode_key()
        B = batch
        f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
        self.layers = spec.layers
        self.A_all = f32(torch.cat([P[f"{l.name}.conv.modulation.weight"] for l in spec.layers], 0))
        self.b_all = f32(torch.cat([P[f"{l.name}.conv.modulation.bias"] for l in spec.layers], 0))
        self.row_widx = torch.cat([torch.full((l.cin,), l.w_idx, dtype=torch.int32) for l in spec.layers]).to(device)
        self.layer_row_start = torch.tensor([l.s_off for l in spec.layers] + [spec.s_dim], dtype=torch.int32, device=device)
        self.layer_widx = torch.tensor([l.w_idx for l in spec.layers], dtype=torch.int32, device=device)
        self.aff_scale = 1.0 / math.sqrt(spec.style_dim)
        c0 = spec.channels[4]
        const = P["input.input"][0].permute(1, 2, 0).contiguous()                      # (4,4,C)
        self.const = const[None].repeat(B, 1, 1, 1).to(device=device, dtype=ACT()).contiguous()
        self.L: List[dict] = []
        max_w = 0
        max_T = 0
        for l in spec.layers:
            e = dict(l=l)
            if l.kind == "rgb":
                e["wrgb"] = f32(P[f"{l.name}.conv.weight"][0, :, :, 0, 0] / math.sqrt(l.cin))   # (3,cin), carries 1/sqrt(cin)
                e["bias"] = f32(P[f"{l.name}.bias"].reshape(3))
                e["rgb"] = _empty((B, 3, l.res, l.res), device, torch.float32)
                e["grgb"] = _empty((B, 3, l.res, l.res), device, torch.float32)
            else:
                W = P[f"{l.name}.conv.weight"][0].to(torch.float32)                              # (cout,cin,3,3)
                scale = 1.0 / math.sqrt(l.cin * 9)
                Ws = (W * scale).to(device)
                e["fused_up"] = l.kind == "up" and l.res >= _fused_up_min_res() and l.cout % 16 == 0 and l.cin % 16 == 0
                if e["fused_up"]:
                    # upsample conv = blur o transposed conv collapsed into four 3x3 phase convs over the input grid
                    # (params.fused_up_base_weights): one launch each way, no phase-plane round trip through HBM
                    We = fused_up_base_weights(Ws).reshape(3, 3, 4 * l.cout, l.cin)
                    e["wbase"] = We.reshape(9, 4 * l.cout, l.cin).contiguous()                    # fp32 [tap][phase*cout][cin]
                    e["wT"] = We.permute(0, 1, 3, 2).reshape(9 * l.cin, 4 * l.cout).to(ACT()).contiguous()
                else:
                    e["wbase"] = Ws.permute(2, 3, 0, 1).reshape(9, l.cout, l.cin).contiguous()        # fp32 [tap][cout][cin]
                    e["wT"] = Ws.permute(2, 3, 1, 0).reshape(9 * l.cin, l.cout).to(ACT()).contiguous()  # [tap][cin][cout], shared
                e["Q"] = (Ws * Ws).sum((2, 3)).contiguous()                                       # (cout,cin)
                e["bias"] = f32(P[f"{l.name}.activate.bias"])
                e["bias_ep"] = (e["bias"] * math.sqrt(2.0)).contiguous()      # for the gain-folded conv epilogue
                e["noise"] = f32(P[f"noises.noise_{l.noise_idx}"][0, 0])
                e["noise_w"] = float(P[f"{l.name}.noise.weight"].reshape(-1)[0])
                e["d"] = _empty((B, l.cout), device, torch.float32)
                e["gdacc"] = _zeros((B, l.cout), device)
                e["out"] = _empty((B, l.res, l.res, l.cout), device)
                e["gout"] = _empty((B, l.res, l.res, l.cout), device)
                max_w = max(max_w, 9 * l.cout * l.cin * (4 if e["fused_up"] else 1))
                if l.kind == "up" and not e["fused_up"]:
                    hp = l.res // 2 + 1
                    max_T = max(max_T, 4 * hp * hp * l.cout)
            self.L.append(e)
        self._pack_style_space(B, device)
        self.T = _empty((B * max(max_T, 8),), device)
        self.gx_scratch = _empty((B, 4, 4, c0), device)
        self.s = _empty((B, spec.s_dim), device, torch.float32)
        self.gs = _zeros((B, spec.s_dim), device)
        self._build_descs()

    def _pack_style_space(self, B, device):
        """Concatenate the per-layer demodulation / weight-modulation operands so that ONE launch each serves every layer
        (sfk_demod_fwd_batched, sfk_modulate_weights_batched, sfk_demod_bwd_batched); the per-layer tensors become views."""
        convs = [e for e in self.L if e["l"].kind != "rgb"]
        q_off = d_off = wb_off = wm_off = 0
        rows_tab = []
        for e in convs:
            l = e["l"]
            rows = e["wbase"].shape[0] * e["wbase"].shape[1]
            # the blur kernel of an unfused up-layer applies d itself; everywhere else d and the activation gain sqrt(2) ride on the
            # weights (lrelu(x)*g == lrelu(g*x)): the conv epilogue is then  max(u, 0.2u),  u = acc + sqrt2*(noise + bias)
            fold = 0 if (l.kind == "up" and not e["fused_up"]) else 2
            rows_tab.append([l.s_off, l.cin, l.cout, q_off, d_off, rows, wb_off, wm_off, l.cout, fold])
            q_off += l.cout * l.cin
            d_off += B * l.cout
            wb_off += rows * l.cin
            wm_off += rows * l.cin
        self.q_cat = _empty((q_off,), device, torch.float32)
        self.d_cat = _empty((d_off,), device, torch.float32)
        self.gd_cat = _zeros((d_off,), device)
        self.wb_cat = _empty((wb_off,), device, torch.float32)
        self.wm_cat = _empty((B * wm_off,), device)
        for e, r in zip(convs, rows_tab):
            l = e["l"]
            _, cin, cout, qo, do, rows, wbo, wmo, _, _ = r
            q = self.q_cat[qo:qo + cout * cin].view(cout, cin)
            q.copy_(e["Q"])
            e["Q"] = q
            wb = self.wb_cat[wbo:wbo + rows * cin].view(e["wbase"].shape)
            wb.copy_(e["wbase"])
            e["wbase"] = wb
            e["d"] = self.d_cat[do:do + B * cout].view(B, cout)
            e["gdacc"] = self.gd_cat[do:do + B * cout].view(B, cout)
            e["wmod"] = self.wm_cat[B * wmo:B * (wmo + rows * cin)].view(B, rows, cin)
        self.style_tab = torch.tensor(rows_tab, dtype=torch.int64, device=device)
        assert self.style_tab.shape[1] == lib.STYLE_TAB_COLS
        self.max_cin = max(e["l"].cin for e in convs)
        self.max_cout = max(e["l"].cout for e in convs)

    # ---------------------------------------------------------------------------------------
    def _build_descs(self):
        B, sd = self.B, self.spec.s_dim
        x, xh = self.const, 4
        prev_conv = None
        for e in self.L:
            l = e["l"]
            if l.kind == "rgb":
                e["x"] = x
                continue
            e["x"] = x
            if l.kind == "conv":
                wmod = e["wmod"]
                e["fwd"] = lib.make_igemm_desc(
                    x, B, l.res, l.res, l.cin, 1, wmod, B, 9 * l.cout, e["out"], l.res, l.res, l.cout, 1, lib.pick_block_n(l.cout),
                    lib.conv3x3_taps(l.cout), flags=lib.EP_NOISE | lib.EP_BIAS | lib.EP_LRELU_RAW,   # demod and gain ride on wmod
                    bias=e["bias_ep"], noise=e["noise"], noise_w=e["noise_w"] * math.sqrt(2.0), err=self.err)
                gx_dst = prev_conv["gout"] if prev_conv is not None else self.gx_scratch
                # The style modulation (x s) and style gradient (sum x*gx~) of a data gradient are finished by the kernel that
                # consumes it (act_bwd / act_torgb_bwd stream both tensors anyway) wherever such a consumer exists; the launch
                # itself then has the plain epilogue.  Exception: conv1 (no consumer).  (The consumer of a conv fed by an unfused
                # up-layer is the blur backward, which finishes the gradient the same way.)
                e["deferred"] = prev_conv is not None
                if e["deferred"]:
                    e["bwd"] = lib.make_igemm_desc(
                        e["gout"], B, l.res, l.res, l.cout, 1, e["wT"], 1, 9 * l.cin, gx_dst, l.res, l.res, l.cin, 1,
                        lib.pick_block_n(l.cin), lib.conv3x3_dgrad_taps(l.cin), err=self.err)
                else:
                    e["bwd"] = lib.make_igemm_desc(
                        e["gout"], B, l.res, l.res, l.cout, 1, e["wT"], 1, 9 * l.cin, gx_dst, l.res, l.res, l.cin, 1,
                        lib.pick_block_n(l.cin), lib.conv3x3_dgrad_taps(l.cin), flags=lib.EP_GSDOT | lib.EP_COLSCALE, xin=x,
                        colscale=self.s, gs=self.gs, vec_stride=sd, vec_off=l.s_off, err=self.err)
            elif e["fused_up"]:
                h = l.res // 2
                wmod = e["wmod"]
                e["fwd"] = lib.make_igemm_desc(
                    x, B, h, h, l.cin, 1, wmod, B, 36 * l.cout, e["out"], h, h, 4 * l.cout, 1, lib.pick_block_n(4 * l.cout),
                    lib.conv3x3_taps(4 * l.cout), flags=lib.EP_NOISE | lib.EP_BIAS | lib.EP_LRELU_RAW,
                    bias=e["bias_ep"], noise=e["noise"], noise_w=e["noise_w"] * math.sqrt(2.0), err=self.err, out_d2s=1)
                e["bwd"] = lib.make_igemm_desc(
                    e["gout"], B, h, h, 4 * l.cout, 1, e["wT"], 1, 9 * l.cin, prev_conv["gout"], h, h, l.cin, 1,
                    lib.pick_block_n(l.cin), lib.conv3x3_dgrad_taps(l.cin), err=self.err, a_s2d=1)   # finished by act_torgb_bwd
                e["deferred"] = True
            else:  # up
                h = l.res // 2
                wmod = e["wmod"]
                T = self.T[: B * 4 * (h + 1) * (h + 1) * l.cout].view(B, 4, h + 1, h + 1, l.cout)
                e["T"] = T
                e["fwd"] = lib.make_igemm_desc(x, B, h, h, l.cin, 1, wmod, B, 9 * l.cout, T, h + 1, h + 1, l.cout, 4,
                                               lib.pick_block_n(l.cout, 4), lib.tconv_taps(l.cout), err=self.err)
                # data gradient is the first writer of the gradient buffer of the previous resolution's conv
                e["bwd"] = lib.make_igemm_desc(
                    T, B, h + 1, h + 1, l.cout, 4, e["wT"], 1, 9 * l.cin, prev_conv["gout"], h, h, l.cin, 1, lib.pick_block_n(l.cin),
                    lib.tconv_dgrad_taps(l.cin), err=self.err)                                         # finished by act_torgb_bwd
                e["deferred"] = True
            x = e["out"]
            prev_conv = e

    # ---------------------------------------------------------------------------------------
    @_on_device
    def styles_from_wplus(self, wplus: torch.Tensor, s_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        s_out = self.s if s_out is None else s_out
        lib.style_affine_fwd(wplus, self.A_all, self.b_all, self.row_widx, s_out, self.aff_scale)
        return s_out

    @_on_device
    def wplus_grad_from_styles(self, gs: torch.Tensor, gw: torch.Tensor) -> torch.Tensor:
        lib.style_affine_bwd(gs, self.A_all, self.layer_row_start, self.layer_widx, gw, self.aff_scale)
        return gw

    @_on_device
    def forward(self, s: Optional[torch.Tensor] = None) -> torch.Tensor:
        """s (B, s_dim) fp32 (default: self.s).  Returns the image buffer (B,3,size,size) fp32."""
        assert self.mode == lib.mode_key(), "engine was built under another activation dtype / conv math (rebuild it)"
        if s is not None and s.data_ptr() != self.s.data_ptr():
            self.s.copy_(s)
        s = self.s
        skip = None
        # every layer's demodulation and per-sample weights in two launches; demodulation rides on the weights wherever the
        # conv epilogue would apply it (the blur kernel of the unfused up-layers applies it itself, after the FIR)
        lib.demod_fwd_batched(s, self.q_cat, self.d_cat, self.style_tab, self.max_cout)
        # (writing the later layers' weights on a side stream while the 4x4 .. 16x16 convs run was measured: 7.81 vs 7.78 ms -- the
        #  weight stream competes with those L2-bound launches for the same bandwidth; one launch on the main stream stays)
        lib.modulate_weights_batched(self.wb_cat, s, self.wm_cat, self.d_cat, self.style_tab)
        for e in self.L:
            l = e["l"]
            if l.kind == "rgb":
                lib.torgb_fwd(e["x"], e["wrgb"], s, l.s_off, e["bias"], skip, e["rgb"])
                skip = e["rgb"]
                continue
            lib.igemm(e["fwd"])
            if l.kind == "up" and not e["fused_up"]:
                lib.blur_act_fwd(e["T"], e["out"], e["d"], e["noise"], e["noise_w"], e["bias"])
        return skip

    @property
    def image(self) -> torch.Tensor:
        return self.L[-1]["rgb"]

    @_on_device
    def weight_grads(self) -> Dict[str, torch.Tensor]:
        """dL/dW (cout,cin,3,3) fp32 of every non-upsampling 3x3 ModulatedConv2d (conv1 and the second conv of each resolution),
        from the buffers forward() + backward() left behind: x (unmodulated input), gout = d * dL/dy, gdacc, d, s -- per-sample
        tensor-core GEMM + the modulation/demodulation chain rule (sfk_modconv_wgrad_finish).  Up-sampling convs and ToRGB are not
        covered (their gradient buffers are scratch that later layers reuse)."""
        res = {}
        for e in self.L:
            l = e["l"]
            if l.kind != "conv":
                continue
            G = lib.conv3x3_wgrad(e["x"], e["gout"], per_sample=True, err=self.err)
            dwb = lib.modconv_wgrad_finish(G, e["wbase"], self.s, l.s_off, e["d"], e["gdacc"])
            scale = 1.0 / math.sqrt(l.cin * 9)
            res[l.name] = (dwb * scale).view(3, 3, l.cout, l.cin).permute(2, 3, 0, 1).contiguous()
        return res

    @_on_device
    def backward(self, g_img: torch.Tensor) -> torch.Tensor:
        """g_img (B,3,size,size) fp32 -> gs (B, s_dim) fp32 (owned).  Must follow forward() with the same s.

        Layer list is [conv1, rgb1, (up, conv, rgb) per resolution].  Every conv output feeds ToRGB and (except at the
        top) the next up-conv: the up-conv's data gradient is the first writer of the conv's gradient buffer, then ONE
        kernel adds ToRGB's backward and applies the conv's activation backward in place (sfk_act_torgb_bwd)."""
        s = self.s
        self.gs.zero_()
        self.gd_cat.zero_()
        L = self.L
        nb = (len(L) - 2) // 3

        def conv_tail(e):
            lib.igemm(e["bwd"])                       # -> gradient of the producer of x (+ style gradient unless deferred)

        def conv_act_rgb(conv, rgb, grgb, producer):
            # ToRGB backward and the conv's activation backward in one pass over conv["out"] / conv["gout"]; `producer` is the
            # up-layer whose plain data gradient already sits in conv["gout"] (its modulation + style gradient are finished here)
            lib.act_torgb_bwd(conv["out"], conv["gout"] if producer is not None else None, conv["gout"], conv["d"], conv["noise"],
                              conv["noise_w"], conv["bias"], conv["gdacc"], rgb["wrgb"], s, rgb["l"].s_off, grgb, self.gs,
                              in_off=producer["l"].s_off if producer is not None else None)

        conv, rgb = (L[3 + 3 * (nb - 1)], L[4 + 3 * (nb - 1)]) if nb > 0 else (L[0], L[1])
        grgb = g_img
        conv_act_rgb(conv, rgb, grgb, None)
        for k in range(nb - 1, -1, -1):
            up, conv = L[2 + 3 * k], L[3 + 3 * k]
            below_conv, below_rgb = (L[3 + 3 * (k - 1)], L[4 + 3 * (k - 1)]) if k > 0 else (L[0], L[1])
            conv_tail(conv)                            # writes up["gout"]
            if up["fused_up"]:
                lib.act_bwd(up["out"], up["gout"], up["gout"], up["d"], up["noise"], up["noise_w"], up["bias"], up["gdacc"],
                            s_in=s, gs_in=self.gs, in_off=conv["l"].s_off)
            else:
                lib.blur_act_bwd(up["out"], up["gout"], up["T"], up["d"], up["noise"], up["noise_w"], up["bias"], up["gdacc"],
                                 s_in=s, gs_in=self.gs, in_off=conv["l"].s_off)
            lib.rgb_down(grgb, below_rgb["grgb"])
            grgb = below_rgb["grgb"]
            conv_tail(up)                              # first writer of below_conv["gout"]
            conv_act_rgb(below_conv, below_rgb, grgb, up)
        conv_tail(L[0])
        # demodulation gradient of every layer (needs each layer's finished gdacc; nothing upstream reads gs before this)
        lib.demod_bwd_batched(s, self.q_cat, self.d_cat, self.gd_cat, self.gs, self.style_tab, self.max_cin)
        return self.gs


# =================================================================================================
@dataclass
class LossCfg:
    c_pix: float = 1.0
    c_feat: float = 1.0
    c_reg: float = 0.0


class AttackEngine:
    """One data-parallel shard of the attack: B independent pairs resident in HBM."""

    def __init__(self, gspec: GenSpec, GP, espec: EncSpec, EP, vgg_sd, FP=None, fusion: str = "arithmetic", batch: int = 1,
                 device="cuda:0", loss: Optional[LossCfg] = None, vgg_res: int = 256, vgg_width_div: int = 1,
                 encoder_module: Optional[torch.nn.Module] = None, latent_avg: Optional[torch.Tensor] = None,
                 n_inputs: int = 2, hierarchy: Optional[dict] = None):
        """n_inputs: images fused into one output (2 = the pairs of BASELINE.json; 5 / 4 / 3 = the reference's ffhq / car / church
        fusion, attack_main2.py:521-581).  fusion: "arithmetic" (mean of the N inputs' W+ codes, interpolation.py:661), "spatial"
        (pair gate) or "hierarchy" (N-way: `hierarchy` = dict(parts=[part names in hierarchy order], source=[input index
        per part], gates={part: dict(alpha, beta, c)}), the s_dict of generate_img (style_fusion_simple.py:84-104) blended by the
        chain of per-part gates that stands in for base_blender.forward (:164); see StyleFusionSimple.hierarchy_for()).
        Rows [k*B:(k+1)*B] of every per-input buffer (x, x0, xin, g_xin, codes ...) belong to input k.
        encoder_module: any torch module `x (n,3,R,R) in [-1,1] -> codes (n, n_latent, 512)` (or (n,512)) that takes the place of the
        encoder stand-in on the gradient path -- e.g. the reference's `net.encoder` = e4e `Encoder4Editing(50,'ir_se')`
        (code/utils/model_utils.py:24; un-vendored upstream).  Its forward and backward run through torch autograd on
        the engine's stream; everything downstream (fusion, synthesis, VGG, update) stays on the CUDA schedules.  `latent_avg`
        (n_latent,512) is added to its codes as get_latents does (attack_main2.py:137-146).  EP (the stand-in's weights) is not
        read then and may be None; CUDA-graph replay is off for such an engine (autograd allocates)."""
        lib.load()
        self.dev = torch.device(device)
        self.gspec, self.espec, self.fusion, self.B = gspec, espec, fusion, batch
        self.loss_cfg = loss or LossCfg()
        self.NI = NI = int(n_inputs)
        assert NI >= 2 and (fusion != "spatial" or NI == 2) and (fusion != "hierarchy" or hierarchy is not None)
        B, dev = batch, self.dev
        S, R = gspec.size, espec.in_res
        self.S, self.R, self.k_in = S, R, S // R
        self.vgg_res, self.k_vgg = vgg_res, S // vgg_res
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        # encoder
        self.enc_module = encoder_module
        self.graph_ok = encoder_module is None
        if encoder_module is None:
            enc_w = [(EP[f"convs.{i}.weight"], EP[f"convs.{i}.bias"]) for i in range(len(espec.widths))]
            self.enc = ConvStack(encoder_layers(espec), enc_w, NI * B, R, dev, self.err)
            self.head_w = f32(EP["head.weight"])
            self.head_b = f32(EP["head.bias"] + EP["latent_avg"].reshape(-1))     # get_latents adds latent_avg (attack_main2.py:137-146)
        else:
            self.enc = None
            self.enc_latent_avg = f32(latent_avg) if latent_avg is not None else None
            for p_ in encoder_module.parameters():      # the attack differentiates w.r.t. the pixels only
                p_.requires_grad_(False)
        cl = espec.widths[-1]
        LD = espec.n_latent * espec.style_dim
        self.feat = _empty((NI * B, cl), dev, torch.float32)
        self.gfeat = _empty((NI * B, cl), dev, torch.float32)
        self.codes = _empty((NI * B, espec.n_latent, espec.style_dim), dev, torch.float32)
        self.gcodes = _empty((NI * B, espec.n_latent, espec.style_dim), dev, torch.float32)
        self.w = _empty((B, espec.n_latent, espec.style_dim), dev, torch.float32)
        self.gw = _empty((B, espec.n_latent, espec.style_dim), dev, torch.float32)
        # synthesis
        self.syn = SynthesisEngine(gspec, GP, B, dev, self.err)
        if fusion == "spatial":
            self.FP = {k: f32(v) for k, v in FP.items()}
        if fusion in ("spatial", "hierarchy"):
            self.s_all = _empty((NI * B, gspec.s_dim), dev, torch.float32)
            self.gs_all = _empty((NI * B, gspec.s_dim), dev, torch.float32)
        if fusion == "hierarchy":
            # out = s[source of parts[0]]; every later part whose input differs from what `out` holds gates it in:
            # out <- q*out + (1-q)*s_part, q = sigmoid(alpha*out + beta*s_part + c)  (oracle/fusion_ref.py blend()).  A part assigned
            # to the base input is skipped only while nothing has been gated in yet (its style vector then EQUALS out).
            parts, source = list(hierarchy["parts"]), [int(k) for k in hierarchy["source"]]
            assert len(parts) == len(source) and all(0 <= k < NI for k in source)
            self.h_base = source[0]
            self.h_chain = []
            for p_, k in zip(parts[1:], source[1:]):
                if not self.h_chain and k == self.h_base:
                    continue
                gate = hierarchy["gates"][p_]
                self.h_chain.append(dict(part=p_, src=k, alpha=f32(gate["alpha"]), beta=f32(gate["beta"]), c=f32(gate["c"])))
            nch = len(self.h_chain)
            self.h_mid = [_empty((B, gspec.s_dim), dev, torch.float32) for _ in range(max(nch - 1, 0))]   # outputs of gates 0..n-2
            self.h_ga = [_empty((B, gspec.s_dim), dev, torch.float32) for _ in range(2)]
            self.h_gb = _empty((B, gspec.s_dim), dev, torch.float32)
        # loss network
        from .params import VGG_EXECUTED
        vals = list(vgg_sd.values())
        vgg_w = [(vals[2 * i], vals[2 * i + 1]) for i in range(VGG_EXECUTED)]
        self.vgg = ConvStack(vgg_layers(vgg_width_div), vgg_w, B, vgg_res, dev, self.err)
        self.vgg_in = _empty((B, 3, vgg_res, vgg_res), dev, torch.float32)
        self.ref_img = _empty((B, 3, S, S), dev, torch.float32)
        self.ref_feats = [torch.empty_like(t) for t in self.vgg.tap_outputs()]
        self.g_img = _empty((B, 3, S, S), dev, torch.float32)
        self.loss = _zeros((B,), dev)
        if self.loss_cfg.c_reg != 0.0:
            self.vgg_reg = ConvStack(vgg_layers(vgg_width_div), vgg_w, NI * B, R, dev, self.err)
            self.reg_refs = [torch.empty_like(t) for t in self.vgg_reg.tap_outputs()]
            self.reg_loss = _zeros((NI * B,), dev)
        # attack state
        self.x = _empty((NI * B, 3, S, S), dev, torch.float32)
        self.x0 = _empty((NI * B, 3, S, S), dev, torch.float32)
        self.xin = _empty((NI * B, 3, R, R), dev, torch.float32)
        self.g_xin = _empty((NI * B, 3, R, R), dev, torch.float32)
        self.stats = _zeros((NI * B,), dev)

    # ---------------------------------------------------------------------------------------
    @_on_device
    def set_inputs(self, *xs: torch.Tensor):
        """the N inputs, each (B,3,S,S) in [0,1] (a single list / tuple of them is accepted too)"""
        if len(xs) == 1 and isinstance(xs[0], (list, tuple)):
            xs = tuple(xs[0])
        B = self.B
        assert len(xs) == self.NI, f"engine fuses {self.NI} inputs, got {len(xs)}"
        for k, xk in enumerate(xs):
            self.x0[k * B:(k + 1) * B].copy_(xk)
        self.x.copy_(self.x0)

    def _rows(self, t: torch.Tensor, k: int) -> torch.Tensor:
        return t[k * self.B:(k + 1) * self.B]

    def _encode(self):
        lib.avgpool_affine_fwd(self.x, self.xin, self.k_in, 2.0, -1.0)
        if self.enc_module is not None:
            self._xin_leaf = self.xin.detach().requires_grad_(True)
            with torch.enable_grad():
                codes = self.enc_module(self._xin_leaf)
            if codes.ndim == 2:                                               # (n,512): one w for every layer (style_fusion_simple.py:139-141)
                codes = codes.unsqueeze(1).expand(-1, self.espec.n_latent, -1)
            self._codes_t = codes
            c = codes.detach().to(torch.float32)
            self.codes.copy_(c + self.enc_latent_avg[None] if self.enc_latent_avg is not None else c)
            return
        top = self.enc.forward(self.xin)
        lib.gap_fwd(top, self.feat)
        lib.linear_fwd(self.feat, self.head_w, self.head_b, self.codes.view(self.NI * self.B, -1))

    def _fuse(self):
        B, NI = self.B, self.NI
        if self.fusion == "arithmetic":
            r = 1.0 / NI
            lib.axpby(self._rows(self.codes, 0), self._rows(self.codes, 1), self.w, r, r)       # interpolation.py:661 (mean over the inputs)
            for k in range(2, NI):
                lib.axpby(self.w, self._rows(self.codes, k), self.w, 1.0, r)
            self.syn.styles_from_wplus(self.w)
        elif self.fusion == "spatial":
            self.syn.styles_from_wplus(self.codes, self.s_all)
            lib.fuse_spatial_fwd(self.s_all[:B], self.s_all[B:], self.FP["alpha"], self.FP["beta"], self.FP["c"], self.syn.s)
        else:
            self.syn.styles_from_wplus(self.codes, self.s_all)
            cur = self._rows(self.s_all, self.h_base)
            n = len(self.h_chain)
            for j, gte in enumerate(self.h_chain):
                dst = self.syn.s if j == n - 1 else self.h_mid[j]
                gte["in"] = cur
                lib.fuse_spatial_fwd(cur, self._rows(self.s_all, gte["src"]), gte["alpha"], gte["beta"], gte["c"], dst)
                cur = dst
            if n == 0:
                self.syn.s.copy_(cur)

    @_on_device
    def fused_forward(self) -> torch.Tensor:
        """current x -> fused image (B,3,S,S) fp32 (buffer owned by the synthesis engine)."""
        self._encode()
        self._fuse()
        return self.syn.forward()

    def _vgg_forward_on_image(self, img):
        lib.avgpool_affine_fwd(img, self.vgg_in, self.k_vgg, 1.0, 0.0)
        self.vgg.forward(self.vgg_in)

    @_on_device
    def compute_reference(self, target: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """Reference = fusion of the clean pair (untargeted) or of `target` (targeted)."""
        B = self.B
        keep = self.x.clone()
        if target is not None:
            for k in range(self.NI):
                self._rows(self.x, k).copy_(target[k])
        else:
            self.x.copy_(self.x0)
        img = self.fused_forward()
        self.ref_img.copy_(img)
        self._vgg_forward_on_image(img)
        for r, t in zip(self.ref_feats, self.vgg.tap_outputs()):
            r.copy_(t)
        if self.loss_cfg.c_reg != 0.0:
            self.x.copy_(self.x0)
            lib.avgpool_affine_fwd(self.x, self.xin, self.k_in, 2.0, -1.0)
            self.vgg_reg.forward(self.xin)
            for r, t in zip(self.reg_refs, self.vgg_reg.tap_outputs()):
                r.copy_(t)
        self.x.copy_(keep)

    @_on_device
    def forward_backward(self):
        """loss (B,) and the gradient w.r.t. the pooled, [-1,1]-mapped inputs g_xin (2B,3,R,R);
        d loss / d x(full res, [0,1]) = (2/k^2) * g_xin[h/k][w/k]."""
        B, cfg = self.B, self.loss_cfg
        S = self.S
        img = self.fused_forward()
        self.loss.zero_()
        # loss + its gradient w.r.t. the fused image
        self._vgg_forward_on_image(img)
        g_vin = self.vgg.backward(self.ref_feats, cfg.c_feat, self.loss) if cfg.c_feat != 0.0 else None
        per = 3 * S * S
        lib.image_loss_grad(img, self.ref_img, g_vin, self.g_img, self.loss, cfg.c_pix / per, 2.0 * cfg.c_pix / per, self.k_vgg)
        # synthesis backward -> style gradient
        gs = self.syn.backward(self.g_img)
        # fusion backward -> latent gradients of both inputs
        NI = self.NI
        if self.fusion == "arithmetic":
            self.syn.wplus_grad_from_styles(gs, self.gw)
            for k in range(NI):
                lib.axpby(self.gw, None, self._rows(self.gcodes, k), 1.0 / NI)
        elif self.fusion == "spatial":
            lib.fuse_spatial_bwd(self.s_all[:B], self.s_all[B:], self.FP["alpha"], self.FP["beta"], self.FP["c"], gs,
                                 self.gs_all[:B], self.gs_all[B:])
            self.syn.wplus_grad_from_styles(self.gs_all, self.gcodes)
        else:
            # the gate chain backwards: the gradient of each gate's second operand accumulates on its input's style vector, the
            # gradient of the first operand walks on down the chain and ends on the base input
            self.gs_all.zero_()
            g = gs
            for j in range(len(self.h_chain) - 1, -1, -1):
                gte = self.h_chain[j]
                ga = self.h_ga[j % 2]
                lib.fuse_spatial_bwd(gte["in"], self._rows(self.s_all, gte["src"]), gte["alpha"], gte["beta"], gte["c"], g, ga, self.h_gb)
                dst = self._rows(self.gs_all, gte["src"])
                lib.axpby(dst, self.h_gb, dst, 1.0, 1.0)
                g = ga
            dst = self._rows(self.gs_all, self.h_base)
            lib.axpby(dst, g, dst, 1.0, 1.0)
            self.syn.wplus_grad_from_styles(self.gs_all, self.gcodes)
        # encoder backward
        if self.enc_module is not None:
            (g_xin,) = torch.autograd.grad(self._codes_t, self._xin_leaf, self.gcodes.to(self._codes_t.dtype))
            g_xin = g_xin.to(torch.float32).contiguous()
            self._codes_t = None
        else:
            lib.linear_bwd(self.gcodes.view(NI * B, -1), self.head_w, self.gfeat)
            lib.gap_bwd(self.enc.out[-1], self.gfeat, self.enc.g[-1])
            g_xin = self.enc.backward(top_grad_ready=True)
        if cfg.c_reg != 0.0:
            # perceptual regulariser on the adversarial inputs themselves (config 5): L -= c_reg * sum_taps MSE
            self.reg_loss.zero_()
            self.vgg_reg.forward(self.xin)
            g_reg = self.vgg_reg.backward(self.reg_refs, -cfg.c_reg, self.reg_loss)
            lib.axpby(g_xin, g_reg, self.g_xin, 1.0, 1.0)
            for k in range(NI):
                lib.axpby(self.loss, self._rows(self.reg_loss, k), self.loss, 1.0, 1.0)
        else:
            self.g_xin.copy_(g_xin)
        return self.loss, self.g_xin

    def full_res_grad(self) -> torch.Tensor:
        """d loss / d x at full resolution (2B,3,S,S) -- diagnostic / parity helper (not used by the loop)."""
        k = self.k_in
        g = self.g_xin.repeat_interleave(k, 2).repeat_interleave(k, 3) if k > 1 else self.g_xin.clone()
        return g * (2.0 / (k * k))

    # ---------------------------------------------------------------------------------------
    def check(self):
        torch.cuda.synchronize(self.dev)
        if int(self.err.item()) != 0:
            raise lib.SfkError("a tensor-core kernel reported an internal pipeline timeout")


# =================================================================================================
@dataclass
class ReconLossCfg:
    """The loss menu of the reference's live loop `optimize_vgg` (code/attack/attack_main2.py:626-649):
    10*l_latent_target + l_img_rec_target - l_latent_org + 20*l_img_org + l_lpips_img  (variants: interpolation.py:818,
    inter_copy.py:658 use other weights and a VGG term on the reconstruction)."""
    w_latent_target: float = 10.0
    w_latent_org: float = -1.0
    w_img_rec_target: float = 1.0
    w_img_org: float = 20.0
    w_lpips_img: float = 1.0
    w_lpips_rec: float = 0.0          # VGG term on the reconstruction ...
    lpips_rec_ref: str = "target"     # ... against the target's ("target") or the clean image's ("org") features


class ReconAttackEngine:
    """Reference-actual attack graph: pixels -> avg-pool -> encoder -> decoder reconstruction of ONE image, losses in latent,
    pixel and VGG space, Adam on the pixels (attack_main2.py:584-671); also serves patch.attack (adversarial_patch.py:94-160).
    Images are in the reference's [-1,1] range."""

    def __init__(self, gspec: GenSpec, GP, espec: EncSpec, EP, vgg_sd, batch: int = 1, device="cuda:0",
                 loss: Optional[ReconLossCfg] = None, vgg_res: int = 256, vgg_width_div: int = 1,
                 encoder_module: Optional[torch.nn.Module] = None):
        """encoder_module: `Model.encoder` as any torch module (x (n,3,R,R) in [-1,1] -> codes (n, n_latent, 512) or (n,512)); its
        forward / backward then run through autograd, the rest of the loop on the CUDA schedules (espec supplies n_latent,
        style_dim and in_res; EP may be None; no CUDA-graph replay)."""
        lib.load()
        self.dev = torch.device(device)
        self.cfg = loss or ReconLossCfg()
        self.B, self.S, self.R = batch, gspec.size, espec.in_res
        self.k_in, self.k_vgg, self.vgg_res = self.S // self.R, self.S // vgg_res, vgg_res
        assert vgg_res == self.R, "the reference pools the image once and feeds both the encoder and VGG (attack_main2.py:619-624)"
        B, dev, S, R = batch, self.dev, self.S, self.R
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        self.enc_module, self.espec = encoder_module, espec
        self.graph_ok = encoder_module is None
        if encoder_module is None:
            enc_w = [(EP[f"convs.{i}.weight"], EP[f"convs.{i}.bias"]) for i in range(len(espec.widths))]
            self.enc = ConvStack(encoder_layers(espec), enc_w, B, R, dev, self.err)
            self.head_w, self.head_b = f32(EP["head.weight"]), f32(EP["head.bias"])     # optimize_vgg calls encoder() directly: no latent_avg
        else:
            self.enc = None
            for p_ in encoder_module.parameters():
                p_.requires_grad_(False)
        L, D, cl = espec.n_latent, espec.style_dim, espec.widths[-1]
        self.LD = L * D
        self.feat, self.gfeat = _empty((B, cl), dev, torch.float32), _empty((B, cl), dev, torch.float32)
        self.codes, self.gcodes, self.gw = (_empty((B, L, D), dev, torch.float32) for _ in range(3))
        self.lat_t, self.lat_o = _empty((B, L, D), dev, torch.float32), _empty((B, L, D), dev, torch.float32)
        self.syn = SynthesisEngine(gspec, GP, B, dev, self.err)
        from .params import VGG_EXECUTED
        vals = list(vgg_sd.values())
        vgg_w = [(vals[2 * i], vals[2 * i + 1]) for i in range(VGG_EXECUTED)]
        self.vgg_img = ConvStack(vgg_layers(vgg_width_div), vgg_w, B, vgg_res, dev, self.err)
        self.vgg_rec = ConvStack(vgg_layers(vgg_width_div), vgg_w, B, vgg_res, dev, self.err) if self.cfg.w_lpips_rec != 0.0 else None
        self.org_feats = [torch.empty_like(t) for t in self.vgg_img.tap_outputs()]
        self.tgt_feats = [torch.empty_like(t) for t in self.vgg_img.tap_outputs()]
        self.x, self.x_org, self.x_tgt, self.g_full, self.g_rec = (_empty((B, 3, S, S), dev, torch.float32) for _ in range(5))
        self.xin, self.g_xin, self.rec_in = (_empty((B, 3, R, R), dev, torch.float32) for _ in range(3))
        self.loss = _zeros((B,), dev)
        # the three terms the reference logs every 5 iterations (attack_main2.py:660-666), unweighted, per sample
        self.terms = _zeros((3, B), dev)      # l_latent_target, l_latent_org, l_img_org
        self.m, self.v = torch.zeros_like(self.x), torch.zeros_like(self.x)

    def _encode(self, x):
        lib.avgpool_affine_fwd(x, self.xin, self.k_in, 1.0, 0.0)
        if self.enc_module is not None:
            self._xin_leaf = self.xin.detach().requires_grad_(True)
            with torch.enable_grad():
                codes = self.enc_module(self._xin_leaf)
            if codes.ndim == 2:
                codes = codes.unsqueeze(1).expand(-1, self.espec.n_latent, -1)
            self._codes_t = codes
            self.codes.copy_(codes.detach().to(torch.float32))
            return
        lib.gap_fwd(self.enc.forward(self.xin), self.feat)
        lib.linear_fwd(self.feat, self.head_w, self.head_b, self.codes.view(self.B, -1))

    @_on_device
    def set_inputs(self, img: torch.Tensor, img_target: torch.Tensor):
        """no_grad setup of optimize_vgg (attack_main2.py:588-603): latents and VGG features of the clean and target image."""
        self.x_org.copy_(img)
        self.x_tgt.copy_(img_target)
        self.x.copy_(img)
        for src, lat, feats in ((self.x_tgt, self.lat_t, self.tgt_feats), (self.x_org, self.lat_o, self.org_feats)):
            self._encode(src)
            lat.copy_(self.codes)
            self.vgg_img.forward(self.xin)
            for r, t in zip(feats, self.vgg_img.tap_outputs()):
                r.copy_(t)
        self.m.zero_()
        self.v.zero_()

    @_on_device
    def reconstruct(self, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._encode(self.x if x is None else x)
        self.syn.styles_from_wplus(self.codes)
        return self.syn.forward()

    @_on_device
    def forward_backward(self):
        """-> loss (B,), g_xin (pooled-resolution gradient; full-res = g_xin[h/k][w/k]/k^2), g_full (direct full-res term).
        self.terms holds the unweighted l_latent_target / l_latent_org / l_img_org of this iteration (the reference's log line)."""
        c, B, S = self.cfg, self.B, self.S
        per = 3 * S * S
        self.loss.zero_()
        self.terms.zero_()
        img_rec = self.reconstruct()
        lib.mse_f32(self.codes, self.lat_t, self.gcodes, self.terms[0], 1.0 / self.LD, 2 * c.w_latent_target / self.LD, False)
        lib.mse_f32(self.codes, self.lat_o, self.gcodes, self.terms[1], 1.0 / self.LD, 2 * c.w_latent_org / self.LD, True)
        lib.axpby(self.loss, self.terms[0], self.loss, 1.0, c.w_latent_target)
        lib.axpby(self.loss, self.terms[1], self.loss, 1.0, c.w_latent_org)
        g_vin = None
        if self.vgg_rec is not None:
            lib.avgpool_affine_fwd(img_rec, self.rec_in, self.k_vgg, 1.0, 0.0)
            self.vgg_rec.forward(self.rec_in)
            refs = self.tgt_feats if c.lpips_rec_ref == "target" else self.org_feats
            g_vin = self.vgg_rec.backward(refs, c.w_lpips_rec, self.loss)
        lib.image_loss_grad(img_rec, self.x_tgt, g_vin, self.g_rec, self.loss, c.w_img_rec_target / per, 2 * c.w_img_rec_target / per,
                            self.k_vgg)
        gs = self.syn.backward(self.g_rec)
        self.syn.wplus_grad_from_styles(gs, self.gw)
        lib.axpby(self.gcodes, self.gw, self.gcodes, 1.0, 1.0)
        if self.enc_module is not None:
            (g_enc,) = torch.autograd.grad(self._codes_t, self._xin_leaf, self.gcodes.to(self._codes_t.dtype))
            g_enc = g_enc.to(torch.float32).contiguous()
            self._codes_t = None
        else:
            lib.linear_bwd(self.gcodes.view(B, -1), self.head_w, self.gfeat)
            lib.gap_bwd(self.enc.out[-1], self.gfeat, self.enc.g[-1])
            g_enc = self.enc.backward(top_grad_ready=True)
        if c.w_lpips_img != 0.0:
            self.vgg_img.forward(self.xin)
            g_v = self.vgg_img.backward(self.org_feats, c.w_lpips_img, self.loss)
            lib.axpby(g_enc, g_v, self.g_xin, 1.0, 1.0)
        else:
            self.g_xin.copy_(g_enc)
        lib.image_loss_grad(self.x, self.x_org, None, self.g_full, self.terms[2], 1.0 / per, 2 * c.w_img_org / per, 1)
        lib.axpby(self.loss, self.terms[2], self.loss, 1.0, c.w_img_org)
        return self.loss, self.g_xin, self.g_full

    @_on_device
    def adam_step(self, t: int, lr: float):
        k = self.k_in
        lib.attack_update_adam(self.x, self.g_xin, self.m, self.v, lr, t, 1.0 / (k * k), k, gfull=self.g_full, gfull_scale=1.0)

    def full_res_grad(self) -> torch.Tensor:
        k = self.k_in
        g = self.g_xin.repeat_interleave(k, 2).repeat_interleave(k, 3) if k > 1 else self.g_xin
        return g / (k * k) + self.g_full

    def check(self):
        torch.cuda.synchronize(self.dev)
        if int(self.err.item()) != 0:
            raise lib.SfkError("a tensor-core kernel reported an internal pipeline timeout")
