#!/usr/bin/env python
"""ncu --csv --metrics ... log (one row per kernel launch and metric) -> a table per kernel: launches, time, warp instructions, issue
utilisation, DRAM bytes.   python tools/ncu_table.py gpurun_out/x.csv [--last-pass N]"""
import collections, csv, sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r and "Metric Name" in r:
            hdr = r; continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        k = (d["ID"], d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("gzb::", ""))
        per.setdefault(k, {})[d["Metric Name"]] = (v, d["Metric Unit"])
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, dict(n=0, ns=0.0, inst=0.0, issue=0.0, rd=0.0, wr=0.0, regs=0, warps=0.0))
        a["n"] += 1
        t, u = m.get("gpu__time_duration.sum", (0, "ns")); a["ns"] += t * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
        a["inst"] += m.get("smsp__inst_executed.sum", (0, ""))[0]
        a["issue"] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
        a["warps"] += m.get("sm__warps_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
        for key, f in (("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum")):
            v, u = m.get(f, (0, "byte")); a[key] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        a["regs"] = int(m.get("launch__registers_per_thread", (0, ""))[0])
    tot_t = sum(a["ns"] for a in agg.values()) or 1; tot_i = sum(a["inst"] for a in agg.values()) or 1
    print(f"# per kernel, all launches of the run ({sys.argv[1]}); times are ncu's serialised, cold-cache times\n")
    print("| kernel | launches | time ms | share | warp instructions | share | issue active % | warps active % | DRAM read MB | DRAM write MB | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda x: -x[1]["inst"]):
        print(f"| `{name}` | {a['n']} | {a['ns'] / 1e6:.2f} | {100 * a['ns'] / tot_t:.1f}% | {a['inst']:.3e} | {100 * a['inst'] / tot_i:.1f}% | {a['issue'] / a['n']:.1f} | {a['warps'] / a['n']:.1f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {a['regs']} |")


if __name__ == "__main__":
    main()
