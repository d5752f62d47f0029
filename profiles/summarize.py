"""Turn the raw ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.
Usage: python profiles/summarize.py   (reads gpurun_out/launches_r1.csv and gpurun_out/prof_spmv_win_r1.ncu-rep)"""
import collections
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_list():
    rows = list(csv.reader(open(os.path.join(G, "launches_r1.csv"))))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            data.append(dict(zip(hdr, r)))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void cosmo::", "").replace("void ", "")
        v, u = float(d["Metric Value"]), d["Metric Unit"]
        ns = v * 1e3 if u in ("usecond", "us") else v * 1e6 if u in ("msecond", "ms") else v
        tot[name][0] += 1
        tot[name][1] += ns
    allns = sum(v[1] for v in tot.values())
    lines = ["# ncu launch list, round 1 (bench.py --steps 3 --warmup 3, config C2, 1 x B200)", "",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1500 --csv "
             "python bench.py --steps 3 --warmup 3 --no-cpu-baseline`",
             "(per-launch times are cold-cache and serialised: compare SHARES, not absolutes; %d launches captured; "
             "the CG iterations run eagerly under ncu's serialisation, from CUDA graphs in the live run)" % len(data), "",
             "| kernel | launches | total ms | share | mean us |", "|---|---:|---:|---:|---:|"]
    for k, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k[:90], c, ns / 1e6, 100 * ns / allns, ns / 1e3 / c))
    open(os.path.join(P, "launches_r1_summary.md"), "w").write("\n".join(lines) + "\n")
    import shutil
    shutil.copy(os.path.join(G, "launches_r1.csv"), os.path.join(P, "launches_r1.csv"))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full_capture():
    out = subprocess.run(["ncu", "-i", os.path.join(G, "prof_spmv_win_r1.ncu-rep"), "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    md = ["# ncu --set full, spmv_win_kernel (round 1, config C2, 1 x B200)", "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:spmv_win -s 60 -c 2 "
          "python bench.py --steps 2 --warmup 3 --no-cpu-baseline`",
          "(profiler replay: cold cache, serialised; the live CUDA-event numbers are in bench_r1_1gpu.json). "
          "Algorithmic bytes per launch: 602.4 MB (A pass) / 608.6 MB (A' + P pass); the measured DRAM traffic is lower "
          "because the slabs carry 16-bit window-local column indices (10 B/nnz).", ""]
    traffic = None
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        md += ["## " + name[:80], "", "| metric | value | unit |", "|---|---:|---|"]
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = (r[i], units[i])
                md.append("| %s | %s | %s |" % (w, r[i], units[i]))
        md.append("")
        if "EpiScale" in name:
            mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
            traffic = float(rd[0]) * mul[rd[1]] + float(wr[0]) * mul[wr[1]]
    open(os.path.join(P, "ncu_spmv_win_r1.md"), "w").write("\n".join(md) + "\n")
    json.dump({"kernel": "spmv_win_kernel<double,EpiScale>", "dram_bytes_per_launch": traffic,
               "source": "profiles/ncu_spmv_win_r1.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(P, "spmv_traffic.json"), "w"))


if __name__ == "__main__":
    launch_list()
    full_capture()
    print(open(os.path.join(P, "launches_r1_summary.md")).read())
    print(open(os.path.join(P, "ncu_spmv_win_r1.md")).read())
