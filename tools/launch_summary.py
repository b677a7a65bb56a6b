"""Summarise an `ncu --csv` launch list (metrics gpu__time_duration.sum, dram__bytes_*.sum, ...) per kernel:
    python tools/launch_summary.py gpurun_out/launches.csv [nlast [nstates [peak_tflops [peak_gbs]]]]
nlast: use only the last `nlast` launches of the list (default: all); nstates: states every launch
processes (default 2^20). Every figure is the AVERAGE PER LAUNCH of that kernel (n = launches seen)."""
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
    nlast = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if nlast > 0:
        items = items[-nlast:]
    tot = OrderedDict()
    for (_, k), m in items:
        k = k.split("(")[0].replace("void ", "")
        t = tot.setdefault(k, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "regs": 0, "issue": 0.0, "fl": 0.0,
                               "inst": 0.0, "fp64": 0.0, "lanes": 0.0, "warps": 0.0})
        t["fl"] += (m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0) +
                    m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0) +
                    2 * m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0))
        t["inst"] += m.get("smsp__inst_executed.sum", 0)
        t["fp64"] += m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0)
        t["lanes"] += m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0)
        t["warps"] += m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)
        t["n"] += 1
        t["ms"] += m.get("gpu__time_duration.sum", 0) * 1e-6
        t["rd"] += m.get("dram__bytes_read.sum", 0) * 1e-9
        t["wr"] += m.get("dram__bytes_write.sum", 0) * 1e-9
        t["regs"] = int(m.get("launch__registers_per_thread", 0))
        t["issue"] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
    for t in tot.values():          # per-launch averages
        for key in ("ms", "rd", "wr", "fl", "inst"):
            t[key] /= t["n"]
    total = sum(t["ms"] for t in tot.values())
    for k, t in tot.items():
        print(f"{k:28s} n={t['n']} {t['ms']:7.3f} ms ({100*t['ms']/total:4.1f}%)  dram rd {t['rd']:6.2f} wr {t['wr']:6.2f} GB"
              f"  regs {t['regs']:3d}  issue {t['issue']/t['n']:5.1f}%")
    print(f"{'one launch of each':28s}     {total:7.3f} ms")
    if any(t["fl"] for t in tot.values()):
        nst = float(sys.argv[3]) if len(sys.argv) > 3 else 1048576.0
        peak_tf = float(sys.argv[4]) if len(sys.argv) > 4 else 36.6
        peak_gbs = float(sys.argv[5]) if len(sys.argv) > 5 else 6541.8
        print(f"\nper kernel and launch, {int(nst):,} states per launch (DFMA probe {peak_tf} TFLOP/s, HBM copy peak {peak_gbs} GB/s):")
        F = B = 0.0
        for k, t in tot.items():
            by = (t["rd"] + t["wr"]) * 1e9
            F += t["fl"]; B += by
            print(f"{k:28s} {t['fl']/nst:8.0f} fp64 flop/state {by/nst:8.0f} DRAM B/state  {t['fl']/t['ms']*1e-9:5.2f} TFLOP/s "
                  f"({100*t['fl']/t['ms']*1e-9/peak_tf:4.1f}%)  {by/t['ms']*1e-6:6.0f} GB/s ({100*by/t['ms']*1e-6/peak_gbs:4.1f}%)  "
                  f"fp64 pipe {t['fp64']/t['n']:4.1f}%  lanes/inst {t['lanes']/t['n']:4.1f}  warps active {t['warps']/t['n']:4.1f}%  "
                  f"warp inst/state {t['inst']/nst:6.0f}")
        print(f"{'step':28s} {F/nst:8.0f} fp64 flop/state {B/nst:8.0f} DRAM B/state  {F/total*1e-9:5.2f} TFLOP/s "
              f"({100*F/total*1e-9/peak_tf:4.1f}%)  {B/total*1e-6:6.0f} GB/s ({100*B/total*1e-6/peak_gbs:4.1f}%)")
        import json
        print(json.dumps({"flops_per_state": round(F/nst, 1), "dram_bytes_per_state": round(B/nst, 1)}))


if __name__ == "__main__":
    main()
