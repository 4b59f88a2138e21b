"""The iterative attack loops on top of AttackEngine (one shard of independent pairs).

Update rules and where they come from in the reference:
  linf  (FGSM / PGD)  code/attack/interpolation.py:71-96 (random start :74-76, step :92, projection :93, clamp :94)
  patch               code/attack/patch/adversarial_patch.py:106,131-138 ; mask apply code/attack/attack_main2.py:413-433
  adam                code/attack/attack_main2.py:606,614-653 (torch.optim.Adam on the pixels)
  l2                  not in the reference: normalised step, projection onto the eps ball
Every step is one fused update kernel (sign/step/projection/clamp/mask + a warp-shuffle reduction); the loss of every
iteration stays in a device buffer and is copied out once after the loop (the reference formats it on the host every
iteration: adversarial_patch.py:141-156).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import lib
from .engine import AttackEngine, LossCfg


@dataclass
class AttackCfg:
    kind: str = "linf"            # linf | l2 | patch | adam
    steps: int = 10
    eps: float = 8.0 / 255.0
    alpha: float = 2.0 / 255.0
    random_start: bool = True
    targeted: bool = False
    patch_sign: bool = False
    lr: float = 1.0
    graph: bool = False           # replay iterations 2..K from one captured CUDA graph (small batches are launch-bound: ~140
                                  # launches per iteration); not for adam (its bias correction is a per-iteration host scalar)


def run_attack(eng: AttackEngine, xa, xb: Optional[torch.Tensor], cfg: AttackCfg, start_noise: Optional[torch.Tensor] = None,
               target: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, mask: Optional[torch.Tensor] = None,
               patch0: Optional[torch.Tensor] = None, compute_final: bool = True, record: Optional[list] = None,
               seed: Optional[int] = None):
    """xa, xb: (B,3,S,S) in [0,1] on the engine's device (N-way fusion: xa = list of the engine's n_inputs tensors, xb = None;
    x_adv then holds input k in rows [k*B:(k+1)*B]).  Returns dict(x_adv, fused_adv, fused_ref, losses, [patch]).
    Random start (interpolation.py:74-76): from `start_noise` (U(-1,1) values supplied by the caller: what the parity tests share
    with the oracle) or, when `seed` is given, drawn on the device by sfk_attack_random_start."""
    B, dev = eng.B, eng.dev
    direction = -1.0 if cfg.targeted else 1.0
    NB = eng.x.shape[0]          # images in the shard: n_inputs * B (pairs: 2B)
    if xb is None:               # N-way fusion: xa is the list of the N inputs
        eng.set_inputs(*xa)
    else:
        eng.set_inputs(xa, xb)
    eng.compute_reference(target if cfg.targeted else None)
    k = eng.k_in
    gscale = 2.0 / (k * k)
    losses = torch.zeros(cfg.steps, B, device=dev)
    if cfg.kind == "patch":
        # one patch per attacked image (the reference runs at batch 1 with a (1,3,S,S) patch, adversarial_patch.py:94-106); a
        # single initial patch is replicated, the copies then evolve independently (a shared, all-reduced patch is SURVEY 8f-4)
        patch = patch0.to(dev).expand_as(eng.x).clone().contiguous()
        mask = mask.to(dev).expand_as(eng.x).contiguous()
        lo = torch.empty(NB, device=dev)
        hi = torch.empty(NB, device=dev)
        lib.minmax_per_sample(eng.x0, lo, hi)
        # adv_x = (1-mask)*img + mask*patch, clamped to the clean range (adversarial_patch.py:106,137-138): a zero-step update
        zero_g = torch.zeros_like(eng.g_xin)
        lib.attack_update_patch(eng.x, eng.x0, patch, mask, zero_g, 0.0, direction, False, lo, hi, gscale, None, k)
    elif cfg.random_start and start_noise is not None and cfg.kind in ("linf", "l2"):
        eng.x.copy_(torch.clamp(eng.x0 + cfg.eps * start_noise.reshape(eng.x0.shape).to(dev), 0.0, 1.0))   # interpolation.py:74-76
    elif cfg.random_start and seed is not None and cfg.kind in ("linf", "l2"):
        lib.attack_random_start(eng.x, eng.x0, cfg.eps, seed)
    if cfg.kind == "adam":
        m = torch.zeros_like(eng.x)
        v = torch.zeros_like(eng.x)
    if cfg.kind == "l2":
        norms = torch.zeros(NB, device=dev)
        dn = torch.zeros(NB, device=dev)
    def iteration(it):
        loss, g = eng.forward_backward()
        if record is not None:      # diagnostics only (forces clones; never used by the benchmark)
            record.append(dict(loss=loss.clone(), grad=eng.full_res_grad(), x=eng.x.clone(), img=eng.syn.image.clone()))
        if cfg.kind == "linf":
            lib.attack_update_linf(eng.x, eng.x0, g, cfg.alpha, cfg.eps, direction, 0.0, 1.0, eng.stats, k)
        elif cfg.kind == "l2":
            norms.zero_()
            dn.zero_()
            for ph in range(3):
                lib.attack_update_l2(eng.x, eng.x0, g, norms, dn, cfg.alpha, cfg.eps, direction, 0.0, 1.0, ph, k)
        elif cfg.kind == "patch":
            lib.attack_update_patch(eng.x, eng.x0, patch, mask, g, cfg.lr, direction, cfg.patch_sign, lo, hi, gscale, eng.stats, k)
        elif cfg.kind == "adam":
            # Adam minimises; ascent (untargeted) flips the gradient sign
            lib.attack_update_adam(eng.x, g, m, v, cfg.lr, it + 1, -direction * gscale, k)
        else:
            raise ValueError(cfg.kind)

    # every buffer an iteration touches is pre-allocated and every launch goes to the current stream, so the body can be captured
    # once (after an eager first iteration has run every kernel's one-time setup) and replayed.  The Linf body only touches
    # engine-owned buffers, so its graph is cached on the engine and reused by later calls with the same step parameters; the
    # L2 / patch bodies use per-call buffers (norms, patch, mask), so they capture per call and only when the run is long.
    use_graph = cfg.graph and getattr(eng, "graph_ok", True) and record is None and cfg.steps > 2 and (cfg.kind == "linf" or (cfg.kind in ("l2", "patch") and cfg.steps >= 20))
    cache = eng.__dict__.setdefault("_iter_graphs", {}) if cfg.kind == "linf" else {}
    key = (cfg.kind, float(cfg.alpha), float(cfg.eps), direction)
    graph = cache.get(key) if use_graph else None
    for it in range(cfg.steps):
        if use_graph and (it >= 1 or graph is not None):
            if graph is None:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA
                    iteration(it)          # recorded, not executed
                cache[key] = graph
            graph.replay()
        else:
            iteration(it)
        losses[it].copy_(eng.loss)
    out = dict(x_adv=eng.x.clone(), fused_ref=eng.ref_img.clone(), losses=losses)
    if compute_final:
        out["fused_adv"] = eng.fused_forward().clone()
    if cfg.kind == "patch":
        out["patch"] = patch
    eng.check()
    return out


def run_attack_stream(eng: AttackEngine, batches, cfg: AttackCfg, seed: int = 0, out_x=None, out_loss=None, gather_into=None):
    """Attack a sequence of host-resident batches with the copies hidden behind the compute.

    `batches`: sequence of (xa_h, xb_h) PINNED host tensors, each (B,3,S,S) fp32 in [0,1] (B = eng.B).  Returns (out_x, out_loss):
    lists of pinned host tensors, out_x[i] (2B,3,S,S) = adversarial pairs of batch i (rows [:B] = a, [B:] = b), out_loss[i]
    (steps,B) = the loss of every iteration.  While batch i runs on the current stream, batch i+1 is copied host->device and the
    results of batch i-1 device->host on two side streams through double-buffered staging tensors (the per-sample attack is
    unchanged: this is run_attack with the random start drawn on the device from seed + i).  linf / l2 only.
    gather_into: optional DEVICE tensor (len(batches)*2B,3,S,S) that also receives every batch's adversarial pairs (the operand of
    the final all-gather across ranks, parallel.gather_results)."""
    assert cfg.kind in ("linf", "l2"), "run_attack_stream: linf / l2 attacks"
    assert getattr(eng, "NI", 2) == 2, "run_attack_stream: pairs (use run_attack for N-way fusion)"
    B, dev = eng.B, eng.dev
    S = eng.S
    nb = len(batches)
    comp = torch.cuda.current_stream(dev)
    st = eng.__dict__.get("_stream_state")
    if st is None or st["steps"] != cfg.steps:
        st = dict(h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev), steps=cfg.steps,
                  xin=[torch.empty(2 * B, 3, S, S, device=dev) for _ in range(2)],
                  xout=[torch.empty(2 * B, 3, S, S, device=dev) for _ in range(2)],
                  lout=[torch.empty(cfg.steps, B, device=dev) for _ in range(2)])
        eng.__dict__["_stream_state"] = st
    h2d, d2h = st["h2d"], st["d2h"]
    if out_x is None:
        out_x = [torch.empty(2 * B, 3, S, S).pin_memory() for _ in range(nb)]
    if out_loss is None:
        out_loss = [torch.empty(cfg.steps, B).pin_memory() for _ in range(nb)]
    ev_in = [torch.cuda.Event() for _ in range(2)]          # staging buffer holds batch i
    ev_taken = [torch.cuda.Event() for _ in range(2)]       # engine has copied the staging buffer
    ev_res = [torch.cuda.Event() for _ in range(2)]         # result staging holds batch i
    ev_out = [torch.cuda.Event() for _ in range(2)]         # result staging has been copied to the host

    def upload(i):
        b = i % 2
        with torch.cuda.stream(h2d):
            if i >= 2:
                h2d.wait_event(ev_taken[b])
            xa_h, xb_h = batches[i]
            st["xin"][b][:B].copy_(xa_h, non_blocking=True)
            st["xin"][b][B:].copy_(xb_h, non_blocking=True)
            ev_in[b].record(h2d)

    h2d.wait_stream(comp)
    upload(0)
    for i in range(nb):
        b = i % 2
        comp.wait_event(ev_in[b])
        xin = st["xin"][b]
        eng.set_inputs(xin[:B], xin[B:])
        ev_taken[b].record(comp)
        if i + 1 < nb:
            upload(i + 1)
        o = _attack_resident(eng, cfg, seed + i)
        if i >= 2:
            comp.wait_event(ev_out[b])
        st["xout"][b].copy_(eng.x)
        st["lout"][b].copy_(o)
        if gather_into is not None:
            gather_into[i * 2 * B:(i + 1) * 2 * B].copy_(eng.x)
        ev_res[b].record(comp)
        with torch.cuda.stream(d2h):
            d2h.wait_event(ev_res[b])
            out_x[i].copy_(st["xout"][b], non_blocking=True)
            out_loss[i].copy_(st["lout"][b], non_blocking=True)
            ev_out[b].record(d2h)
    comp.wait_stream(d2h)
    eng.check()
    return out_x, out_loss


def _attack_resident(eng: AttackEngine, cfg: AttackCfg, seed: int) -> torch.Tensor:
    """run_attack's body for inputs already placed with eng.set_inputs (untargeted linf / l2): reference fusion, device random
    start, cfg.steps iterations (graph replay as in run_attack).  Returns the (steps, B) loss buffer (engine-cached)."""
    B, dev, k = eng.B, eng.dev, eng.k_in
    eng.compute_reference(None)
    if cfg.random_start:
        lib.attack_random_start(eng.x, eng.x0, cfg.eps, seed)
    key = ("losses", cfg.steps)
    losses = eng.__dict__.setdefault("_loss_bufs", {}).get(key)
    if losses is None:
        losses = torch.zeros(cfg.steps, B, device=dev)
        eng.__dict__["_loss_bufs"][key] = losses
    if cfg.kind == "l2":
        norms = eng.__dict__.setdefault("_l2_norms", torch.zeros(2 * B, device=dev))
        dn = eng.__dict__.setdefault("_l2_dn", torch.zeros(2 * B, device=dev))

    def iteration():
        _, g = eng.forward_backward()
        if cfg.kind == "linf":
            lib.attack_update_linf(eng.x, eng.x0, g, cfg.alpha, cfg.eps, 1.0, 0.0, 1.0, eng.stats, k)
        else:
            norms.zero_()
            dn.zero_()
            for ph in range(3):
                lib.attack_update_l2(eng.x, eng.x0, g, norms, dn, cfg.alpha, cfg.eps, 1.0, 0.0, 1.0, ph, k)

    cache = eng.__dict__.setdefault("_iter_graphs", {})
    gkey = (cfg.kind, float(cfg.alpha), float(cfg.eps), 1.0)
    use_graph = cfg.graph and getattr(eng, "graph_ok", True)
    graph = cache.get(gkey) if use_graph else None
    for it in range(cfg.steps):
        if use_graph and (it >= 1 or graph is not None):
            if graph is None:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    iteration()
                cache[gkey] = graph
            graph.replay()
        else:
            iteration()
        losses[it].copy_(eng.loss)
    return losses
