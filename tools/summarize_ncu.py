#!/usr/bin/env python
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches_rNN.csv profiles/rNN_launches.md
  python tools/summarize_ncu.py kernel   gpurun_out/prof_X.ncu-rep   profiles/rNN_X.md [kernel-name substring]
"""
import collections, csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if "Kernel Name" in r:
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                agg[d["Kernel Name"].split("(")[0]].append(float(d["Metric Value"].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum per kernel, --clock-control none\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | max ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            if sum(v) / tot < 0.0005:
                continue
            f.write(f"| `{k[:70]}` | {len(v)} | {sum(v)/1e6:.2f} | {max(v)/1e6:.2f} | {100*sum(v)/tot:.1f}% |\n")


def kernel(src, dst, name=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    if name:                                               # last launch whose kernel name contains `name`
        for r in rows[2:]:
            if name in r[h.index("Kernel Name")]:
                v = r
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n| metric | value | unit |\n|---|---:|---|\n")
        name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        f.write(f"| kernel | `{name[:90]}` | |\n")
        for m in WANT:
            if m in h:
                i = h.index(m)
                f.write(f"| {m} | {v[i]} | {units[i]} |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:])
