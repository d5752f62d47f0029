"""Turn the raw round-2 ncu outputs (gpurun_out/*.ncu-rep, launches_r2_*.csv; see profiles/capture_r2.sh) into the
committed summaries under profiles/.  Usage: python profiles/summarize_r2.py"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
if len(sys.argv) > 1:          # on the GPU box: write the summaries next to the raw files (only gpurun_out/ travels back)
    P = sys.argv[1]
    os.makedirs(P, exist_ok=True)
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}

METRICS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB"),
    ("dram__bytes_write.sum", "MB"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64 %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        d["_units"] = dict(zip(hdr, units))
        res.append(d)
    return res


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    return v * m.get(unit, 1)


def short(name):
    name = name.replace("void ", "").replace("cosmo::", "").replace("tc::", "")
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("(int)", "")
    return name[:110]


def summarize(rep, title, note=""):
    rows = raw_rows(rep)
    if not rows:
        return None
    agg = collections.OrderedDict()
    for d in rows:
        k = short(d.get("Kernel Name", "?"))
        agg.setdefault(k, []).append(d)
    lines = ["# %s" % title, "", "Source: `%s` (`ncu --set full --clock-control none`, see profiles/capture_r2.sh)." % os.path.basename(rep),
             "Peak HBM = %.0f GB/s (MEASURED_PEAKS.json).  Durations are per launch under the profiler (cold cache, serialised)." % PEAKS["hbm_gbs"],
             "", note, "",
             "| kernel | launches | us | DRAM rd+wr MB | DRAM GB/s | of HBM peak | tensor % | fp64 % | warps % | regs | grid x block | L2 % |",
             "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|"]
    for k, ds in agg.items():
        def avg(metric, conv=None):
            vals = []
            for d in ds:
                v = fnum(d.get(metric, ""))
                if v is None:
                    continue
                u = d["_units"].get(metric, "")
                vals.append(conv(v, u) if conv else v)
            return sum(vals) / len(vals) if vals else None
        us = avg("gpu__time_duration.sum", to_us)
        rd = avg("dram__bytes_read.sum", to_bytes) or 0.0
        wr = avg("dram__bytes_write.sum", to_bytes) or 0.0
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us else None
        f = lambda x, fmt="%.1f": (fmt % x) if x is not None else "-"
        lines.append("| `%s` | %d | %s | %s | %s | %s | %s | %s | %s | %s | %s x %s | %s |" % (
            k, len(ds), f(us), f((rd + wr) / 1e6), f(gbs, "%.0f"), f(gbs / PEAKS["hbm_gbs"] if gbs else None, "%.2f"),
            f(avg("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")),
            f(avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed")),
            f(avg("sm__warps_active.avg.pct_of_peak_sustained_active")),
            f(avg("launch__registers_per_thread"), "%.0f"), f(avg("launch__grid_size"), "%.0f"), f(avg("launch__block_size"), "%.0f"),
            f(avg("lts__throughput.avg.pct_of_peak_sustained_elapsed"))))
    return "\n".join(lines) + "\n"


def launch_list(csv_path, title, cmd):
    rows = list(csv.reader(open(csv_path, errors="ignore")))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            data.append(dict(zip(hdr, r)))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        name = short(d["Kernel Name"])
        v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
        tot[name][0] += 1
        tot[name][1] += to_us(v, u)
    allus = sum(v[1] for v in tot.values()) or 1.0
    lines = ["# %s" % title, "", "Command: `%s`" % cmd,
             "(per-launch times are cold-cache and serialised: compare SHARES, not absolutes; %d launches captured)" % len(data), "",
             "| kernel | launches | total ms | share | mean us |", "|---|---:|---:|---:|---:|"]
    for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k, c, us / 1e3, 100 * us / allus, us / c))
    return "\n".join(lines) + "\n"


def main():
    jobs = [
        ("ncu_r2_tc.ncu-rep", "ncu_tc_gemm_r2.md", "Tensor-core product kernel of the PSD projection (N = 2000, 8 slices, 10 groups) and the slicing kernel",
         "`ozaki_gemm_kernel`: bound = tensor pipe (int8 tcgen05.mma); `slice_rows_kernel`: bound = HBM/L2 (reads 8 N^2 bytes, writes 8 N^2 bytes of int8 slices)."),
        ("ncu_r2_c4.ncu-rep", "ncu_c4_vector_kernels_r2.md", "Config C4 (n = 2 001 000, m = 2 003 000): projection / rhs kernel and the elementwise kernels around the PSD projection", ""),
        ("ncu_r2_c2.ncu-rep", "ncu_c2_cg_and_residual_kernels_r2.md", "Config C2: CG vector kernels, windowed SpMV with its epilogues (incl. the residual passes of the first termination check)", ""),
        ("ncu_r2_c5.ncu-rep", "ncu_c5_psd_small_r2.md", "Config C5 (|V| = 3000, parent-child merge): batched shared-memory Jacobi over the clique cones", ""),
        ("ncu_r2_bj.ncu-rep", "ncu_block_jacobi_r2.md", "Block-Jacobi fallback eigensolver (N = 1000), kept for the infeasibility certificates and as the fallback of the tensor-core path", ""),
    ]
    for rep, out, title, note in jobs:
        path = os.path.join(G, rep)
        if not os.path.exists(path):
            print("missing", rep)
            continue
        txt = summarize(path, title, note)
        if txt:
            open(os.path.join(P, out), "w").write(txt)
            print("wrote", out)
    for src, out, title, cmd in [
        ("launches_r2_c2.csv", "launches_r2_c2_summary.md", "ncu launch list, round 2, config C2 (bench.py, 1 x B200)",
         "ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1500 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline"),
        ("launches_r2_c4.csv", "launches_r2_c4_summary.md", "ncu launch list, round 2, config C4 (tests/run_configs.py c4, 1 x B200)",
         "ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv python tests/run_configs.py c4"),
    ]:
        path = os.path.join(G, src)
        if os.path.exists(path):
            open(os.path.join(P, out), "w").write(launch_list(path, title, cmd))
            print("wrote", out)


if __name__ == "__main__":
    main()
