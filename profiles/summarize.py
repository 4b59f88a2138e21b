"""Turn the ncu CSV exports brought back in gpurun_out/ into the summaries committed next to this file.

  python profiles/summarize.py launches gpurun_out/launches.csv  profiles/launches_rN      -> .md + .csv (one PGD step)
  python profiles/summarize.py full     gpurun_out/igemm_full.csv profiles/igemm_full_rN   -> .md, and igemm_traffic_rN.json
  python profiles/summarize.py elem     gpurun_out/elem_full.csv  profiles/elementwise_full_rN -> .md (bandwidth-bound kernels)

`launches`: output of  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...  over a bench.py run; the
step is the launch window between two consecutive update_linf_kernel launches (one PGD iteration of all pairs).
`full`: output of  ncu --set full --clock-control none -k regex:igemm_tc2 -c <launches per step> --csv --page raw.
"""
import collections
import csv
import json
import sys


def _rows(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    return rows[h], rows[h + 1:]


def _short(name):
    return name.replace("void ", "").replace("<unnamed>::", "").split("(")[0][:64]


def launches(src, dst, cmd=""):
    H, data = _rows(src)
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
    names = [r[ki] for r in data]
    us = [float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3) for r in data]
    upd = [i for i, n in enumerate(names) if "update_linf" in n]
    assert len(upd) >= 2, "capture window does not hold a full step (need two update_linf_kernel launches)"
    a, b = upd[-2] + 1, upd[-1] + 1
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in zip(names[a:b], us[a:b]):
        agg[_short(n)][0] += 1
        agg[_short(n)][1] += v
    tot = sum(v for _, v in agg.values())
    with open(dst + ".csv", "w") as f:
        f.write("index,kernel,us\n")
        for i in range(a, b):
            f.write(f"{i - a},\"{_short(names[i])}\",{us[i]:.2f}\n")
    with open(dst + ".md", "w") as f:
        f.write("# ncu launch list of one PGD step (8 pairs, 1024x1024, 1 B200)\n\n")
        if cmd:
            f.write(f"Command: `{cmd}`\n")
        f.write(f"({b - a} consecutive launches between two `update_linf_kernel` launches = one PGD iteration; per-launch times are\n"
                "cold-cache and serialised by the profiler: compare SHARES with bench.py, not absolutes).\n\n")
        f.write(f"total {tot / 1e3:.2f} ms over {b - a} launches\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |\n")
    ig = sum(v for k, (c, v) in agg.items() if "igemm" in k)
    print(f"{b - a} launches, {tot / 1e3:.2f} ms; igemm share {100 * ig / tot:.1f}%")


def full(src, dst, cmd="", shapes_json=""):
    """shapes_json: bench.py --dump-launches output of the same step (its conv_launches are in launch order): adds the GEMM shape,
    the event-timed duration and the algorithmic TFLOP/s of every row (the launch <-> ncu-row map)."""
    H, data = _rows(src)
    shapes = json.load(open(shapes_json))["conv_launches"] if shapes_json else None

    def col(suffix):
        if suffix in H:
            return H.index(suffix)
        for i, h in enumerate(H):
            if h.endswith(suffix):
                return i
        return None

    units = None
    # the raw page has a units row right after the header
    if data and not data[0][0].strip().isdigit():
        units, data = data[0], data[1:]
    ci = {k: col(k) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "Grid Size", "Block Size",
                              "launch__registers_per_thread", "sm__inst_executed_pipe_tensor.sum",
                              "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                              "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed")}

    def val(r, k, unit_scale=None):
        i = ci.get(k)
        if i is None or r[i] in ("", "n/a"):
            return float("nan")
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        if unit_scale and units:
            v *= unit_scale.get(units[i], 1.0)
        return v

    tsc = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}
    bsc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rows = []
    for r in data:
        t = val(r, "gpu__time_duration.sum", tsc)
        by = val(r, "dram__bytes_read.sum", bsc) + val(r, "dram__bytes_write.sum", bsc)
        rows.append((r[ci["Grid Size"]] if ci["Grid Size"] is not None else "", t, by,
                     val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                     val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
                     val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")))
    n = len(rows)
    tot_t = sum(r[1] for r in rows)
    tot_b = sum(r[2] for r in rows)
    regs = val(data[0], "launch__registers_per_thread")
    with open(dst + ".md", "w") as f:
        f.write(f"# ncu --set full: the {n} tensor-core conv launches of one PGD iteration (8 pairs, 1024x1024)\n\n")
        if cmd:
            f.write(f"Command (after the same command exited 0 without ncu): `{cmd}`\n\n")
        f.write(f"{n} launches: total {tot_t / 1e3:.2f} ms (cold-cache, serialised), DRAM traffic {tot_b / 1e9:.2f} GB = "
                f"**{tot_b / n / 1e6:.1f} MB per launch** on average; registers/thread {regs:.0f}.\n\n")
        if shapes and len(shapes) == n:
            wt = sum(r[1] * r[4] for r in rows) / tot_t
            f.write(f"Time-weighted tensor-pipe activity over the {n} launches: **{wt:.1f} %**.  `shape` = images x output grid, Cin->Cout "
                    "(x4 accumulators for the 4-phase transposed conv; fused upsample convs list their 4*Cout columns); `event us` / "
                    "`TFLOP/s` are the CUDA-event timings of the same launches inside an unprofiled step (bench.py --dump-launches).\n\n")
            f.write("| # | shape | grid | ncu us | event us | alg. TFLOP/s | DRAM MB (r+w) | achieved GB/s | DRAM % | tensor pipe % | L2 % |\n"
                    "|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
            for i, (r, c) in enumerate(zip(rows, shapes)):
                shp = f"{c['n']}x{c['out_h']}x{c['out_w']} {c['cin']}->{c['cout']}" + (f" x{c['num_acc']}acc" if c['num_acc'] > 1 else "")
                f.write(f"| {i} | {shp} | {r[0]} | {r[1]:.1f} | {c['us']:.1f} | {c['tflops']:.0f} | {r[2] / 1e6:.1f} | {r[2] / r[1] / 1e3:.0f} | {r[3]:.1f} | "
                        f"{r[4]:.1f} | {r[5]:.1f} |\n")
        else:
            f.write("| # | grid | us | DRAM MB (r+w) | DRAM % | tensor pipe % | L2 % |\n|---:|---|---:|---:|---:|---:|---:|\n")
            for i, r in enumerate(rows):
                f.write(f"| {i} | {r[0]} | {r[1]:.1f} | {r[2] / 1e6:.1f} | {r[3]:.1f} | {r[4]:.1f} | {r[5]:.1f} |\n")
    tr = dst.replace("igemm_full", "igemm_traffic") + ".json"
    json.dump({"kernel": "igemm_tc2_kernel", "launches_per_step": n, "dram_bytes_per_launch": tot_b / n, "dram_bytes_per_step": tot_b,
               "serialized_ms_per_step": tot_t / 1e3, "source": dst + ".md (ncu --set full, the conv launches of one step)"},
              open(tr, "w"), indent=1)
    print(f"{n} launches, {tot_t / 1e3:.2f} ms, {tot_b / 1e9:.2f} GB DRAM -> {tr}")


def elem(src, dst, cmd=""):
    """bandwidth-bound kernels of one step: per launch time, DRAM bytes and achieved GB/s (ncu --set full, raw page)"""
    H, data = _rows(src)
    units = None
    if data and not data[0][0].strip().isdigit():
        units, data = data[0], data[1:]

    def col(name):
        return H.index(name) if name in H else next((i for i, h in enumerate(H) if h.endswith(name)), None)

    ci = {k: col(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                              "dram__throughput.avg.pct_of_peak_sustained_elapsed")}
    tsc = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}
    bsc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, k, sc=None):
        i = ci[k]
        try:
            v = float(r[i].replace(",", ""))
        except (ValueError, TypeError):
            return float("nan")
        return v * sc.get(units[i], 1.0) if (sc and units) else v

    with open(dst + ".md", "w") as f:
        f.write("# ncu --set full: bandwidth-bound kernels of one PGD iteration (8 pairs, 1024x1024)\n\n")
        if cmd:
            f.write(f"Command (after the same command exited 0 without ncu): `{cmd}`\n\n")
        f.write("Achieved GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration of that launch (cold-cache, serialised).\n\n")
        f.write("| # | kernel | us | DRAM MB (r+w) | achieved GB/s | % of the measured copy peak (6547 GB/s) |\n|---:|---|---:|---:|---:|---:|\n")
        for i, r in enumerate(data):
            t = val(r, "gpu__time_duration.sum", tsc)
            by = val(r, "dram__bytes_read.sum", bsc) + val(r, "dram__bytes_write.sum", bsc)
            f.write(f"| {i} | `{_short(r[ci['Kernel Name']])}` | {t:.1f} | {by / 1e6:.1f} | {by / t / 1e3:.0f} | "
                    f"{100 * by / t / 1e3 / 6547:.1f} |\n")
    print(f"{len(data)} launches -> {dst}.md")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    extra = sys.argv[5:6]
    {"launches": launches, "full": full, "elem": elem}[mode](src, dst, cmd, *extra)
