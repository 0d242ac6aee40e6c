#!/usr/bin/env python
"""Turn the ncu artefacts that gpurun brings back (gpurun_out/) into the small text summaries
committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv   > profiles/rNN_launches.txt
  python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep   > profiles/rNN_<kernel>.txt
  python profiles/summarize.py kernel   gpurun_out/X.raw.csv gpurun_out/X.src.csv   (exports made on the GPU box)

`launches`: per-kernel count / total / average device time and SHARE of the step from the
`--metrics gpu__time_duration.sum` pass (cold-cache, serialised: shares are what is comparable).
`kernel`: the metrics the roofline discussion in DESIGN.md quotes, from one `--set full` capture,
plus the warp-stall mix and the hottest SASS lines of the source page.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.3f ms total (gpu__time_duration.sum, cold-cache, serialised)" % (path, len(rows) - 1, tot / 1e6))
    print("%-58s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %5d %12.1f %10.1f %6.1f%%" % (k[:58], v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))


def kernel(path, src_path=None):
    """path: a .ncu-rep, or the `--page raw --csv` export of one (then src_path = its `--page source --csv
    --print-source sass` export: the report itself may have stayed on the GPU box)"""
    if path.endswith(".csv"):
        raw = "\n".join(l for l in open(path).read().splitlines() if l.startswith('"'))
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        if name in seen:
            continue
        seen.add(name)
        print("== %s  grid %s block %s" % (name, d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                print("   %-72s %s %s" % (k, d[k], units[hdr.index(k)]))
        st = {k: float(d[k]) for k in hdr if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and d[k]}
        tot = sum(st.values())
        print("   warp stall mix (cycles per issued instruction, share):")
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
            print("      %-28s %6.2f  %5.1f%%" % (k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v, 100 * v / tot))
    if path.endswith(".csv"):
        if not src_path:
            return
        src = open(src_path).read()
    else:
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    done = set()
    for si, st in enumerate(starts):
        name = rows[st][1]
        if name in done:
            continue
        done.add(name)
        seg = rows[st + 1: starts[si + 1] if si + 1 < len(starts) else None]
        hdr = seg[0]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in seg[1:] if len(r) >= len(hdr) - 2]
        samples = lambda r: int(r[ix["# Samples"]] or 0)
        tot = sum(samples(r) for r in data) or 1
        print("== hottest SASS lines of %s (%d instructions, %d samples)" % (name[:70], len(data), tot))
        for i in sorted(sorted(range(len(data)), key=lambda i: -samples(data[i]))[:12]):
            r = data[i]
            reasons = {k[6:]: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
            big = ", ".join("%s %d" % kv for kv in sorted(reasons.items(), key=lambda kv: -kv[1])[:2] if kv[1])
            print("   %4d  %-58s %6d %5.1f%%  %s" % (i, r[ix["Source"]].strip()[:58], samples(r), 100 * samples(r) / tot, big))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:])
