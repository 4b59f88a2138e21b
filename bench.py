#!/usr/bin/env python
"""Benchmark of the attack hot path (BASELINE.json metric: PGD attack-iterations x images / second).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      the CPU oracle port of the same step (rank 0 only)

A "step" is ONE PGD iteration (encoder -> fusion -> StyleGAN2 synthesis -> VGG+pixel loss -> backward ->
fused sign/projection/clamp update) over this GPU's shard of independent image pairs.
  N = 1: BASELINE.json configs[1] -- PGD L-inf (eps 8/255, alpha 2/255, random start), StyleGAN2-1024 (config-f, random-init),
         arithmetic (mean W+) fusion, 8 pairs resident, bf16 activations / fp32 accumulate.
  N > 1: BASELINE.json configs[2] -- the same attack on spatial (StyleSpace gate) fusion, global batch 64 sharded contiguously over
         the ranks (64/N pairs per GPU, attacked as resident batches of 8); strong scaling, with the weak-scaled number (8 pairs per
         GPU) in `weak_scaled`.  Pairs are independent: no collective inside the loop; one NCCL all-gather of the adversarial examples
         after it (inside the end-to-end region).
`value` is device-timed with the shard resident in HBM; `e2e` goes through attack_loop.run_attack_stream from pinned host buffers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pgd_attack_iterations_x_images_per_sec"
UNIT = "iter*img/s"
SIZE = 1024
PAIRS_PER_GPU = 8
EPS, ALPHA = 8.0 / 255.0, 2.0 / 255.0


def synthetic_pairs(n_pairs: int, size: int, first_index: int = 0):
    """U[0,1] images smoothed by a 5x5 box blur, seed 1234 + pair index; generated on the host."""
    import torch.nn.functional as F
    xs = []
    for i in range(n_pairs):
        g = torch.Generator().manual_seed(1234 + first_index + i)
        x = torch.rand(2, 3, size + 4, size + 4, generator=g)
        xs.append(F.avg_pool2d(x, 5, 1))
    x = torch.stack(xs)                       # (n, 2, 3, S, S)
    return x[:, 0].contiguous(), x[:, 1].contiguous()


def start_noise(n_pairs: int, size: int, first_index: int = 0):
    g = torch.Generator().manual_seed(4321 + first_index)
    return torch.rand(2, n_pairs, 3, size, size, generator=g) * 2 - 1


def build_models(size: int):
    from sfattack.params import EncSpec, gen_spec, make_encoder_params, make_generator_params, make_vgg_state_dict
    spec = gen_spec(size)
    GP = make_generator_params(spec, seed=0)
    es = EncSpec(n_latent=spec.n_latent)
    EP = make_encoder_params(es, seed=1)
    vsd = make_vgg_state_dict(2)
    return spec, GP, es, EP, vsd


def flops_per_iter_image(spec, es) -> dict:
    """Algorithmic FLOPs of one attack iteration on one pair: forward + data-gradient (2x forward) of the generator
    and of VGG (360.3 GFLOP at 1024), plus the encoder stand-in on the pair's two inputs (reported apart)."""
    g = 0.0
    for l in spec.layers:
        hw = l.res * l.res
        if l.kind == "rgb":
            g += 2.0 * hw * l.cin * 3
        elif l.kind == "up":
            g += 2.0 * (hw / 4) * 9 * l.cin * l.cout
        else:
            g += 2.0 * hw * 9 * l.cin * l.cout
    from sfattack.params import VGG_CONVS
    v, r = 0.0, 256
    pools_after = {1, 3, 6}
    for i, (_, ci, co) in enumerate(VGG_CONVS[:9]):
        v += 2.0 * r * r * 9 * ci * co
        if i in pools_after:
            r //= 2
    e, r, ci = 0.0, es.in_res, 3
    for i, co in enumerate(es.widths):
        e += 2.0 * r * r * 9 * ci * co
        ci = co
        if i < len(es.widths) - 1:
            r //= 2
    return dict(generator_fwd=g, vgg_fwd=v, encoder_fwd_pair=2 * e, attack=2.0 * (g + v), with_encoder=2.0 * (g + v + 2 * e))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvidia-smi needs ~0.1 s to start, so it is launched before the warm-up; only rows stamped inside [mark_begin, mark_end]
    # (the timed region) are kept.

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.02)]
        if not rows:
            rows = [r for _, r in self.rows[-3:]]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d.get("bf16_tflops_sustained", 1393.1), hbm_gbs=d.get("hbm_gbs", 6547.2), source="measured")
    return dict(bf16_tflops=1400.0, hbm_gbs=6650.0, source="fallback")


# ---------------------------------------------------------------------------------------------------
def cpu_oracle_iteration_rate(size: int, iters: int, warm: int = 1, fusion: str = "arithmetic"):
    """The oracle port (oracle/pipeline.py) on the host cores: PGD iterations on ONE pair of the same workload."""
    from oracle.pipeline import LossCfg as OLoss, OraclePipeline, linf_step
    torch.set_num_threads(os.cpu_count() or 1)
    spec, GP, es, EP, vsd = build_models(size)
    from sfattack.params import make_fusion_params
    FP = make_fusion_params(spec.s_dim) if fusion == "spatial" else None
    pipe = OraclePipeline(spec, GP, es, EP, vsd, FP, fusion=fusion)
    xa, xb = synthetic_pairs(1, size)
    nz = start_noise(1, size)
    with torch.no_grad():
        ref_img, ref_feats = pipe.reference_of(pipe.fused(xa, xb))
    X0 = torch.cat([xa, xb])
    X = torch.clamp(X0 + EPS * nz.reshape(X0.shape), 0, 1)
    times = []
    for it in range(warm + iters):
        t0 = time.perf_counter()
        L, img, ga, gb = pipe.input_grads(X[:1], X[1:], ref_img, ref_feats, OLoss(1.0, 1.0))
        X = linf_step(X, X0, torch.cat([ga, gb]), ALPHA, EPS, 1.0)
        if it >= warm:
            times.append(time.perf_counter() - t0)
    return len(times) / sum(times), times


GLOBAL_PAIRS_C3 = 64   # BASELINE.json configs[2]: batch 64 sharded across 2/4/8 GPUs


def workload_config(steps, size, pairs, world):
    """the `config` both arms print: the same workload, whoever executes it.  N = 1: BASELINE.json configs[1]; N > 1: configs[2]."""
    loss = "pixel+VGG(conv1_1,conv1_2,pool2,conv4_2) loss at 256x256, encoder stand-in on the gradient path"
    l2 = "working set per step (>2 GB of activations) exceeds the 126 MB L2; no explicit flush"
    if world == 1:
        return {"workload": f"PGD-{steps} Linf eps=8/255 alpha=2/255 random-start, StyleGAN2-{size} config-f random-init, "
                            f"arithmetic (mean W+) fusion of pairs, {loss}; BASELINE.json configs[1]",
                "pairs_per_gpu": pairs, "global_pairs": pairs, "image_size": size, "attack": "linf-pgd", "fusion": "arithmetic",
                "parallelism": "dp1 (independent pairs, no in-loop collective)", "l2_policy": l2}
    per = GLOBAL_PAIRS_C3 // world
    return {"workload": f"PGD-{steps} Linf eps=8/255 alpha=2/255 random-start, StyleGAN2-{size} config-f random-init, "
                        f"spatial (StyleSpace gate) fusion of pairs, {loss}; BASELINE.json configs[2]: global batch {GLOBAL_PAIRS_C3} "
                        f"sharded over {world} GPUs ({per} pairs per GPU, attacked in resident batches of {pairs})",
            "pairs_per_gpu": per, "global_pairs": GLOBAL_PAIRS_C3, "image_size": size, "attack": "linf-pgd", "fusion": "spatial",
            "parallelism": f"dp{world} (independent pairs sharded contiguously, no in-loop collective; one NCCL all-gather of the "
                           f"adversarial examples after the loop, inside the e2e region)", "l2_policy": l2}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    world = max(1, args.gpus)
    fusion = "arithmetic" if world == 1 else "spatial"
    rate, times = cpu_oracle_iteration_rate(SIZE, args.steps, args.warmup, fusion=fusion)
    cores = os.cpu_count() or 1
    sample = (f"{args.steps} PGD iterations on 1 pair of the workload at {SIZE}x{SIZE}, {fusion} fusion "
              f"(fp32 PyTorch CPU oracle port, params frozen)")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.steps, SIZE, PAIRS_PER_GPU, world),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from sfattack import lib
    from sfattack.attack_loop import AttackCfg, _attack_resident, run_attack_stream
    from sfattack.engine import AttackEngine, LossCfg
    from sfattack.params import make_fusion_params
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path is mandatory; use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    B = args.pairs
    S = args.size
    spec, GP, es, EP, vsd = build_models(S)
    # N = 1: configs[1] (arithmetic fusion, one resident batch of 8 pairs).  N > 1: configs[2] (spatial fusion, global batch 64
    # sharded contiguously: 64/N pairs per rank, attacked as 64/(N*8) resident batches of 8 pairs)
    fusion = "arithmetic" if world == 1 else "spatial"
    n_local = B if world == 1 else max(B, GLOBAL_PAIRS_C3 // world)
    n_batches = n_local // B
    FP = make_fusion_params(spec.s_dim) if fusion == "spatial" else None
    eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion=fusion, batch=B, device=str(dev), loss=LossCfg(1.0, 1.0))
    first = rank * n_local
    host = []
    for b in range(n_batches):       # this rank's shard of the pair range, in pinned host memory
        xa_h, xb_h = synthetic_pairs(B, S, first_index=first + b * B)
        host.append((xa_h.pin_memory(), xb_h.pin_memory()))
    resident = [(xa_h.to(dev, non_blocking=True), xb_h.to(dev, non_blocking=True)) for xa_h, xb_h in host]
    cfg = AttackCfg(kind="linf", steps=args.steps, eps=EPS, alpha=ALPHA, random_start=True, graph=args.graph)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # -------- device-resident timing: every pair of the shard already in HBM; K PGD iterations on each resident batch
    k = eng.k_in
    eng.set_inputs(*resident[0])
    eng.compute_reference()
    lib.attack_random_start(eng.x, eng.x0, EPS, 4321 + first)

    def step():
        _, g = eng.forward_backward()
        lib.attack_update_linf(eng.x, eng.x0, g, ALPHA, EPS, 1.0, 0.0, 1.0, eng.stats, k)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    eng.check()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_graph = None
    if os.environ.get("SFK_NCU_RANGE") == "1":
        args.graph = False                                # profile plain launches
        cfg.graph = False
    lib.LAUNCHES = 0
    if args.graph:                                        # same launches, recorded once (after the warm-up ran them eagerly)
        step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph, capture_error_mode="thread_local"):
            step()
        step_graph.replay()                               # one untimed replay: the GPU is busy again when the clock window opens
        torch.cuda.synchronize(dev)
    else:
        step()
        torch.cuda.synchronize(dev)
    launches_per_step = lib.LAUNCHES

    def run_step():
        if step_graph is not None:
            step_graph.replay()
        else:
            step()

    barrier()
    sampler.mark_begin()
    profiled = os.environ.get("SFK_NCU_RANGE") == "1"     # ncu --profile-from-start off: capture only the timed steps
    if profiled:
        torch.cuda.profiler.start()
    lib.LAUNCHES = 0
    ev0.record()
    if n_batches == 1:
        for _ in range(args.steps):
            run_step()
    else:
        for xa_d, xb_d in resident:                       # per resident batch: place it, reference fusion, random start, K iterations
            eng.set_inputs(xa_d, xb_d)
            eng.compute_reference()
            lib.attack_random_start(eng.x, eng.x0, EPS, 4321 + first)
            for _ in range(args.steps):
                run_step()
    ev1.record()
    if profiled:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    sampler.mark_end()
    launches = lib.LAUNCHES + (launches_per_step * args.steps * n_batches if step_graph is not None else 0)
    clocks = sampler.stop()
    eng.check()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps * n_local * world / (ms * 1e-3)
    step_ms = ms / args.steps                             # one PGD iteration over the rank's whole shard
    weak = None
    if world > 1:       # beside the sharded configs[2] number: the weak-scaled one (one resident batch of B pairs per GPU, K iterations)
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(args.steps):
            run_step()
        w1.record()
        barrier()
        tw = torch.tensor([w0.elapsed_time(w1)], device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak = {"value": args.steps * B * world / (float(tw.item()) * 1e-3), "unit": UNIT, "pairs_per_gpu": B,
                "ms_per_step": float(tw.item()) / args.steps, "scaling": "weak"}

    # -------- end to end through the public call with HOST buffers (attack_loop.run_attack_stream): per resident batch H2D of the
    # pairs from pinned memory, reference fusion, device random start, PGD-K, D2H of adversarial examples + per-iteration losses;
    # the copies of neighbouring batches overlap the compute on side streams.  N > 1: + the final NCCL all-gather.
    e2e_rounds = max(1, args.e2e_calls // n_batches) if n_batches > 1 else 1
    e2e_batches = host * e2e_rounds if n_batches > 1 else host * max(1, args.e2e_calls)
    out_x = [torch.empty(2 * B, 3, S, S).pin_memory() for _ in range(len(e2e_batches))]
    out_l = [torch.empty(args.steps, B).pin_memory() for _ in range(len(e2e_batches))]
    x_local = torch.empty(len(e2e_batches) * 2 * B, 3, S, S, device=dev) if world > 1 else None
    x_all = torch.empty(world * len(e2e_batches) * 2 * B, 3, S, S, device=dev) if world > 1 else None

    def e2e_call():
        run_attack_stream(eng, e2e_batches, cfg, seed=4321 + first, out_x=out_x, out_loss=out_l, gather_into=x_local)
        if world > 1:                                     # the only collective: adversarial examples of every rank, after the loop
            dist.all_gather_into_tensor(x_all, x_local)
        torch.cuda.synchronize(dev)

    e2e_call()  # warm (allocates the staging buffers, captures nothing new)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_call()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = len(e2e_batches) * B * args.steps * world / (e2e_ms * 1e-3)
    steps_total = len(e2e_batches) * args.steps / max(1, n_batches if n_batches > 1 else 1)
    h2d = sum(a.numel() + b.numel() for a, b in e2e_batches) * 4 / steps_total
    d2h = (sum(o.numel() for o in out_x) + sum(o.numel() for o in out_l)) * 4 / steps_total

    # -------- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv): per-launch CUDA-event timing of one step
    peaks = measured_peaks()
    # (three eager steps, per launch the SHORTEST of its three timings: an event pair around an eager launch also counts any time
    #  the stream waited for the host to submit it, and a host hiccup -- the clock sampler's subprocess, a busy box -- once inflated
    #  a single-step sum by a third; the kernel's own duration is the minimum)
    profs = [lib.profile_igemm(step) for _ in range(3)]
    prof = profs[0]
    n_tc, tc_flops = prof["launches"], prof["flops"]
    if all(p["launches"] == n_tc for p in profs):
        tc_ms = sum(min(p["per_launch"][i][0] for p in profs) for i in range(n_tc))
    else:
        tc_ms = min(p["ms"] for p in profs)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    fl = flops_per_iter_image(spec, es)
    batch_step_ms = step_ms / n_batches                   # one PGD iteration on one resident batch of B pairs
    traffic = None
    tp = os.path.join(ROOT, "profiles", "igemm_traffic_r2.json")
    if os.path.exists(tp) and S == SIZE and B == PAIRS_PER_GPU:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")      # ncu --set full, dram read+write per launch (profiles/igemm_full_r2.md)
    roofline = {"bound": "tensor", "kernel": "igemm_tc2_kernel", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic, "peak_source": peaks["source"] + " (sustained bf16)",
                "launches_per_step": n_tc, "kernel_ms_per_step": tc_ms, "kernel_share_of_step": tc_ms / batch_step_ms if batch_step_ms else None,
                "algorithmic_gflop_per_iter_image": fl["attack"] / 1e9,
                "step_frac_of_roofline": (fl["attack"] * B / (batch_step_ms * 1e-3)) / (peaks["bf16_tflops"] * 1e12)}

    # -------- optional diagnostics (never part of a reported number): per-launch conv timings with the planner's decisions, and a
    # per-kernel time table of one eager step from torch's profiler
    if rank == 0 and args.dump_launches:
        rows = [dict(us=1e3 * ms_, gflop=f / 1e9, tflops=(f / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0), n=shp[0], out_h=shp[1], out_w=shp[2],
                     cin=shp[3], cout=shp[4], taps=shp[5], num_acc=shp[6]) for ms_, f, shp in prof["per_launch"]]
        infos = []
        for st in (eng.enc, eng.vgg):
            infos += [lib.plan_info(d) for d in list(st._fwd_desc.values()) + list(st._bwd_desc.values())]
        for e in eng.syn.L:
            infos += [lib.plan_info(e[k_]) for k_ in ("fwd", "bwd") if k_ in e]
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as tp_:
            step()
            torch.cuda.synchronize(dev)
        table = sorted(((ev.key, ev.count, ev.device_time_total) for ev in tp_.key_averages()), key=lambda r: -r[2])
        json.dump(dict(step_ms=batch_step_ms, conv_launches=rows, plans=infos, kernel_table=[dict(name=n_[:120], count=c_, us=u_) for n_, c_, u_ in table]),
                  open(args.dump_launches, "w"), indent=1)

    # -------- N = 1 only: the single-GPU point of the N > 1 workload (configs[2]: spatial fusion, all 64 pairs as 8 resident batches,
    # per-batch reference fusion + random start inside the timed region exactly as the N > 1 arm times it), so that the strong-scaling
    # efficiency of the 2/4/8-GPU lines has its own denominator (their `value` is not comparable with this line's configs[1] `value`)
    c3_single = None
    if world == 1 and not args.no_c3_line and S == SIZE and B == PAIRS_PER_GPU:
        del prof
        eng3 = AttackEngine(spec, GP, es, EP, vsd, make_fusion_params(spec.s_dim), fusion="spatial", batch=B, device=str(dev), loss=LossCfg(1.0, 1.0))
        res3 = []
        for b in range(GLOBAL_PAIRS_C3 // B):
            xa_h, xb_h = synthetic_pairs(B, S, first_index=b * B)
            res3.append((xa_h.to(dev), xb_h.to(dev)))
        k3 = eng3.k_in

        def step3():
            _, g3 = eng3.forward_backward()
            lib.attack_update_linf(eng3.x, eng3.x0, g3, ALPHA, EPS, 1.0, 0.0, 1.0, eng3.stats, k3)

        eng3.set_inputs(*res3[0])
        eng3.compute_reference()
        lib.attack_random_start(eng3.x, eng3.x0, EPS, 4321)
        for _ in range(args.warmup):
            step3()
        g3 = None
        if args.graph:
            g3 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g3, capture_error_mode="thread_local"):
                step3()
            g3.replay()
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for xa_d, xb_d in res3:
            eng3.set_inputs(xa_d, xb_d)
            eng3.compute_reference()
            lib.attack_random_start(eng3.x, eng3.x0, EPS, 4321)
            for _ in range(args.steps):
                g3.replay() if g3 is not None else step3()
        c1.record()
        torch.cuda.synchronize(dev)
        eng3.check()
        c3_ms = c0.elapsed_time(c1)
        c3_single = {"value": args.steps * GLOBAL_PAIRS_C3 / (c3_ms * 1e-3), "unit": UNIT, "global_pairs": GLOBAL_PAIRS_C3, "ms_per_step": c3_ms / args.steps,
                     "what": "BASELINE.json configs[2] on ONE GPU (spatial fusion, 64 pairs as 8 resident batches of 8; per-batch reference "
                             "fusion and random start inside the timed region, as in the N > 1 lines): the denominator of their strong scaling"}
        del eng3, res3

    if rank == 0:
        cpu = None
        if world > 1:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only (rank 0 of the 1-GPU run)"}
        elif not args.no_cpu_baseline:
            rate, times = cpu_oracle_iteration_rate(S, args.cpu_iters, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{args.cpu_iters} PGD iteration(s) on 1 of the {B} pairs at {S}x{S} (oracle/pipeline.py, fp32, "
                             f"{sum(times):.1f}s)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args.steps, S, B, world),
                "clocks": clocks, "gpu_launches": launches, "cuda_graph": bool(args.graph),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms": e2e_ms,
                        "what": f"attack_loop.run_attack_stream over {len(e2e_batches)} batch(es) of {B} pairs from pinned host buffers: "
                                f"H2D pairs, reference fusion, device random start, PGD-{args.steps}, D2H adversarial examples + per-iteration "
                                f"losses (copies of neighbouring batches overlap the compute)"
                                + (", then the NCCL all-gather of all adversarial examples" if world > 1 else "")},
                "roofline": roofline, "cpu_baseline": cpu,
                "per_gpu_iter_img_per_s": value / world, "weak_scaled": weak, "config3_single_gpu": c3_single,
                "scaling_note": None if world == 1 else "strong scaling of configs[2]: compare with config3_single_gpu.value of the N=1 line (the same "
                                "64-pair workload on one GPU); the N=1 headline `value` is configs[1] (arithmetic fusion, no per-batch setup in the "
                                "timed region) and `weak_scaled.value` is its N-GPU counterpart",
                "encoder_gflop_per_iter_image": 2 * fl["encoder_fwd_pair"] / 1e9}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--e2e-calls", type=int, default=4, help="resident batches in the timed end-to-end region (N = 1)")
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3-line", action="store_true", help="N = 1: skip the single-GPU point of the N > 1 workload (configs[2])")
    ap.add_argument("--dump-launches", default=None, help="diagnostics: write per-launch conv timings + a per-kernel time table (JSON)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel eagerly instead of replaying the step / the e2e iterations from captured CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
