#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` (one row per captured launch, one column per metric) -> a markdown table of the metrics that say what
bounds a kernel, and (with --traffic OUT.json VBLOCKS) the dram bytes per launch per kernel for bench.py's roofline.traffic.
    python tools/ncu_raw_table.py gpurun_out/r02_fastq64_raw.csv [--traffic profiles/r02_traffic.json 64]"""
import csv, json, sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("smsp__average_warp_latency_per_inst_issued.ratio", "cyc/inst"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "short sb"), ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "branch"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3, "nsecond": 1e-6}


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    out = ["| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---:|" * len(WANT)]
    traffic = {}
    for r in rows[2:]:
        if len(r) != len(h):
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("gzb::", "").replace("<unnamed>::", "")
        cells = []
        rd = wr = 0.0
        for m, n in WANT:
            if m not in h:
                cells.append("-"); continue
            i = h.index(m)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                cells.append(r[i]); continue
            u = units[i]
            if n == "time":
                v *= UNIT.get(u, 1); cells.append(f"{v:.2f} ms")
            elif n in ("DRAM rd", "DRAM wr"):
                v *= UNIT.get(u, 1)
                if n == "DRAM rd": rd = v
                else: wr = v
                cells.append(f"{v / 1e6:.1f} MB")
            elif n == "warp inst":
                cells.append(f"{v:.3e}")
            elif n in ("grid", "block", "regs"):
                cells.append(str(int(v)))
            else:
                cells.append(f"{v:.2f}")
        out.append(f"| `{name}` | " + " | ".join(cells) + " |")
        t = traffic.setdefault(name, []); t.append(rd + wr)
    print("\n".join(out))
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic"); dst, vb = sys.argv[i + 1], int(sys.argv[i + 2])
        try:
            tj = json.load(open(dst))
        except Exception:
            tj = {}
        for k, v in traffic.items():
            tj[k] = {"vblocks": vb, "bytes_per_launch": int(max(v)), "launches_captured": len(v), "source": sys.argv[1].replace("gpurun_out/", "profiles/"),
                     "note": "dram__bytes_read.sum + dram__bytes_write.sum of the largest captured launch of this kernel (ncu --set full --clock-control none)"}
        json.dump(tj, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
