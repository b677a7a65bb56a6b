"""Summarise an `ncu --csv` launch list (metrics gpu__time_duration.sum, dram__bytes_*.sum, ...)
per kernel over the LAST bench step: python tools/launch_summary.py gpurun_out/launches.csv [nlast]"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    d = OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    items = list(d.items())
    nlast = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    items = items[-nlast:]
    tot = OrderedDict()
    for (_, k), m in items:
        k = k.split("(")[0].replace("void ", "")
        t = tot.setdefault(k, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "regs": 0, "issue": 0.0})
        t["n"] += 1
        t["ms"] += m.get("gpu__time_duration.sum", 0) * 1e-6
        t["rd"] += m.get("dram__bytes_read.sum", 0) * 1e-9
        t["wr"] += m.get("dram__bytes_write.sum", 0) * 1e-9
        t["regs"] = int(m.get("launch__registers_per_thread", 0))
        t["issue"] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
    total = sum(t["ms"] for t in tot.values())
    for k, t in tot.items():
        print(f"{k:28s} n={t['n']} {t['ms']:7.3f} ms ({100*t['ms']/total:4.1f}%)  dram rd {t['rd']:6.2f} wr {t['wr']:6.2f} GB"
              f"  regs {t['regs']:3d}  issue {t['issue']/t['n']:5.1f}%")
    print(f"{'total':28s}     {total:7.3f} ms")


if __name__ == "__main__":
    main()
