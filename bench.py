#!/usr/bin/env python
"""Benchmark of the attack hot path (BASELINE.json metric: PGD attack-iterations x images / second).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      the CPU oracle port of the same step (rank 0 only)

A "step" is ONE PGD iteration (encoder -> fusion -> StyleGAN2 synthesis -> VGG+pixel loss -> backward ->
fused sign/projection/clamp update) over this GPU's batch of independent image pairs.  Workload at every N:
BASELINE.json configs[1] -- PGD L-inf (eps 8/255, alpha 2/255, random start), StyleGAN2-1024 (config-f,
random-init), arithmetic (mean W+) fusion, 8 pairs per GPU, bf16 activations / fp32 accumulate.  Weak scaling:
pairs are independent, each rank owns 8, no collective inside the loop; one NCCL all-gather of the adversarial
examples after the timed region.
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
def cpu_oracle_iteration_rate(size: int, iters: int, warm: int = 1):
    """The oracle port (oracle/pipeline.py) on the host cores: PGD iterations on ONE pair of the same workload."""
    from oracle.pipeline import LossCfg as OLoss, OraclePipeline, linf_step
    torch.set_num_threads(os.cpu_count() or 1)
    spec, GP, es, EP, vsd = build_models(size)
    pipe = OraclePipeline(spec, GP, es, EP, vsd, None, fusion="arithmetic")
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


def workload_config(steps, size, pairs, world):
    """the `config` both arms print: the same workload, whoever executes it"""
    return {"workload": f"PGD-{steps} Linf eps=8/255 alpha=2/255 random-start, StyleGAN2-{size} config-f random-init, "
                        f"arithmetic (mean W+) fusion of pairs, pixel+VGG(conv1_1,conv1_2,pool2,conv4_2) loss at 256x256, "
                        f"encoder stand-in on the gradient path; BASELINE.json configs[1]",
            "pairs_per_gpu": pairs, "global_pairs": pairs * world, "image_size": size, "attack": "linf-pgd",
            "parallelism": f"dp{world} (independent pairs, no in-loop collective)",
            "l2_policy": "working set per step (>2 GB of activations) exceeds the 126 MB L2; no explicit flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    rate, times = cpu_oracle_iteration_rate(SIZE, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = f"{args.steps} PGD iterations on 1 of the {PAIRS_PER_GPU} pairs at {SIZE}x{SIZE} (fp32 PyTorch CPU oracle port, params frozen)"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.steps, SIZE, PAIRS_PER_GPU, max(1, args.gpus)),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from sfattack import lib
    from sfattack.attack_loop import AttackCfg, run_attack
    from sfattack.engine import AttackEngine, LossCfg
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
    spec, GP, es, EP, vsd = build_models(args.size)
    eng = AttackEngine(spec, GP, es, EP, vsd, None, fusion="arithmetic", batch=B, device=str(dev), loss=LossCfg(1.0, 1.0))
    # this rank's shard of the (weak-scaled) pair range, in pinned host memory
    xa_h, xb_h = synthetic_pairs(B, args.size, first_index=rank * B)
    nz_h = start_noise(B, args.size, first_index=rank * B)
    xa_h, xb_h, nz_h = xa_h.pin_memory(), xb_h.pin_memory(), nz_h.pin_memory()
    xa, xb, nz = xa_h.to(dev, non_blocking=True), xb_h.to(dev, non_blocking=True), nz_h.to(dev, non_blocking=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # -------- device-resident timing: inputs already in HBM, K PGD iterations
    eng.set_inputs(xa, xb)
    eng.compute_reference()
    eng.x.copy_(torch.clamp(eng.x0 + EPS * nz.reshape(eng.x0.shape), 0.0, 1.0))
    k = eng.k_in

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
    if args.graph:                                        # same launches, recorded once (after the warm-up ran them eagerly)
        lib.LAUNCHES = 0
        step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph, capture_error_mode="thread_local"):
            step()
        launches_per_step = lib.LAUNCHES
        step_graph.replay()                               # one untimed replay: the GPU is busy again when the clock window opens
        torch.cuda.synchronize(dev)
    barrier()
    sampler.mark_begin()
    lib.LAUNCHES = 0
    profiled = os.environ.get("SFK_NCU_RANGE") == "1"     # ncu --profile-from-start off: capture only the timed steps
    if profiled:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        if step_graph is not None:
            step_graph.replay()
        else:
            step()
    ev1.record()
    if profiled:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    sampler.mark_end()
    launches = lib.LAUNCHES if step_graph is None else launches_per_step * args.steps
    clocks = sampler.stop()
    eng.check()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps * B * world / (ms * 1e-3)

    # -------- end to end through the public call with HOST buffers: H2D pairs, PGD-K, D2H adversarial examples + losses
    cfg = AttackCfg(kind="linf", steps=args.steps, eps=EPS, alpha=ALPHA, random_start=True, graph=args.graph)
    out_h = torch.empty(2 * B, 3, args.size, args.size).pin_memory()
    loss_h = torch.empty(args.steps, B).pin_memory()
    e2e_calls = max(1, args.e2e_calls)

    def e2e_call():
        xa_d = xa_h.to(dev, non_blocking=True)
        xb_d = xb_h.to(dev, non_blocking=True)
        nz_d = nz_h.to(dev, non_blocking=True)
        o = run_attack(eng, xa_d, xb_d, cfg, start_noise=nz_d, compute_final=False)
        out_h.copy_(o["x_adv"], non_blocking=True)
        loss_h.copy_(o["losses"], non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_call()  # warm
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_calls):
        e2e_call()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = e2e_calls * args.steps * B * world / (e2e_ms * 1e-3)
    h2d = (xa_h.numel() + xb_h.numel() + nz_h.numel()) * 4 / args.steps
    d2h = (out_h.numel() + loss_h.numel()) * 4 / args.steps

    # -------- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv): per-launch CUDA-event timing of one step
    peaks = measured_peaks()
    prof = lib.profile_igemm(step)
    tc_flops, tc_ms, n_tc = prof["flops"], prof["ms"], prof["launches"]
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    fl = flops_per_iter_image(spec, es)
    step_ms = ms / args.steps
    traffic = None
    tp = os.path.join(ROOT, "profiles", "igemm_traffic_r1.json")
    if os.path.exists(tp) and args.size == SIZE and B == PAIRS_PER_GPU:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")      # ncu --set full, dram read+write per launch (profiles/igemm_full_r1.md)
    roofline = {"bound": "tensor", "kernel": "igemm_tc2_kernel", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic, "peak_source": peaks["source"] + " (sustained bf16)",
                "launches_per_step": n_tc, "kernel_ms_per_step": tc_ms, "kernel_share_of_step": tc_ms / step_ms if step_ms else None,
                "algorithmic_gflop_per_iter_image": fl["attack"] / 1e9,
                "step_frac_of_roofline": (fl["attack"] * B / (step_ms * 1e-3)) / (peaks["bf16_tflops"] * 1e12)}

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
        json.dump(dict(step_ms=step_ms, conv_launches=rows, plans=infos, kernel_table=[dict(name=n_[:120], count=c_, us=u_) for n_, c_, u_ in table]),
                  open(args.dump_launches, "w"), indent=1)

    # -------- final gather of adversarial examples + metrics (the only collective; outside the loop)
    gather_ms = None
    if world > 1:
        x_all = torch.empty(world * 2 * B, 3, args.size, args.size, device=dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(x_all, eng.x)
        g1.record()
        torch.cuda.synchronize(dev)
        gather_ms = g0.elapsed_time(g1)

    if rank == 0:
        cpu = None
        if world > 1:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only (rank 0 of the 1-GPU run)"}
        elif not args.no_cpu_baseline:
            rate, times = cpu_oracle_iteration_rate(args.size, args.cpu_iters, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{args.cpu_iters} PGD iteration(s) on 1 of the {B} pairs at {args.size}x{args.size} (oracle/pipeline.py, fp32, "
                             f"{sum(times):.1f}s)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": workload_config(args.steps, args.size, B, world),
                "clocks": clocks, "gpu_launches": launches, "cuda_graph": bool(args.graph),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "what": f"{e2e_calls} call(s) of attack_loop.run_attack from pinned host buffers: H2D pairs+start noise, reference "
                                f"fusion, PGD-{args.steps}, D2H adversarial examples + per-iteration losses"},
                "roofline": roofline, "cpu_baseline": cpu, "final_gather_ms": gather_ms,
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
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
