#!/usr/bin/env python
"""Per-source-line cost of one kernel from an ncu report captured with --import-source on (-lineinfo build):
  python tools/ncu_lines.py REPORT.ncu-rep KERNEL LAUNCH_SKIP UNITS
prints, for every source line above 0.3 %, its share of stall samples, of executed warp instructions, and warp
instructions per unit (UNITS = symbols or steps the launch processed)."""
import csv, subprocess, sys


def main():
    rep, kern, skip, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, lines, h = "?", [], None
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            h = r
        elif h and len(r) == len(h) and r[0] not in ("", "Line No"):
            lines.append((fname, r))
    si, ie = h.index("# Samples"), h.index("Instructions Executed")
    tot = sum(int(r[si] or 0) for _, r in lines); toti = sum(int(r[ie] or 0) for _, r in lines)
    print(f"samples {tot}  warp-instructions {toti}  per unit {toti/units:.1f}")
    for f, r in lines:
        s, n = int(r[si] or 0), int(r[ie] or 0)
        if s * 300 > tot or n * 300 > toti:
            print(f"{100*s/tot:5.1f}% smp {100*n/toti:5.1f}% ins {n/units:6.1f}/u  {f}:{r[0]:>4} {r[1][:130]}")


if __name__ == "__main__":
    main()
