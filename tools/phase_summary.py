"""Per-kernel summary of an `ncu --set full` capture of the phase kernels:
    python tools/phase_summary.py gpurun_out/prof.ncu-rep NSTATES [peak_tflops peak_gbs] > profiles/rNN_phase_kernels_ncu.txt
NSTATES = states processed by each captured launch (one chunk). Prints the metrics the roofline in
bench.py is built from and, on the last line, a JSON record for profiles/flops_per_state.json."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    rep, nstates = sys.argv[1], float(sys.argv[2])
    peak_tf = float(sys.argv[3]) if len(sys.argv) > 3 else 36.6
    peak_gbs = float(sys.argv[4]) if len(sys.argv) > 4 else 6541.8
    if rep.endswith(".csv"):      # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x.csv
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    tot = {"flops": 0.0, "bytes": 0.0, "t": 0.0}
    print(f"ncu --set full --clock-control none, one chunk ({int(nstates):,} humanoid states, contacts on, pyramidal, "
          f"qM/qLD/qLDiagInv written) of\npython bench.py --steps 1 --warmup 3 --no-cpu-baseline   on B200\n")
    for r in data:
        name = r[col["Kernel Name"]]
        print("==", name)
        v = {}
        for k in KEYS:
            if k not in col:
                continue
            x = float(r[col[k]].replace(",", "")) if r[col[k]] not in ("", "n/a") else 0.0
            u = units[col[k]]
            v[k] = x * UNIT.get(u, 1.0) if k.startswith(("dram__bytes", "gpu__time")) else x
            print(f"{k:88s} {r[col[k]]:>18s} {u}")
        if "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum" not in v:
            # `--set full` of this ncu carries the fp64 op counts only per cycle: flops per kernel are
            # taken from the launch list instead (tools/launch_summary.py)
            by = v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
            t = v["gpu__time_duration.sum"]
            print(f"derived: {by/nstates:9.0f} DRAM B/state {by/t*1e-9:7.0f} GB/s DRAM "
                  f"({100*by/t*1e-9/peak_gbs:4.1f}% of HBM copy peak)\n")
            continue
        fl = (v["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] +
              v["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] +
              2 * v["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"])
        by = v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
        t = v["gpu__time_duration.sum"]
        tot["flops"] += fl; tot["bytes"] += by; tot["t"] += t
        print(f"derived: {fl/nstates:9.0f} fp64 flop/state {by/nstates:9.0f} DRAM B/state {fl/t*1e-12:6.2f} TFLOP/s "
              f"({100*fl/t*1e-12/peak_tf:4.1f}% of DFMA probe) {by/t*1e-9:7.0f} GB/s DRAM ({100*by/t*1e-9/peak_gbs:4.1f}% of HBM copy peak)\n")
    if tot["t"] == 0:
        return
    print(f"== sum over the kernels: {tot['flops']/nstates:.0f} fp64 flop/state, {tot['bytes']/nstates:.0f} DRAM B/state, "
          f"{tot['t']*1e3:.2f} ms per chunk ({nstates/tot['t']:.3g} states/s under ncu), "
          f"{tot['flops']/tot['t']*1e-12:.2f} TFLOP/s = {100*tot['flops']/tot['t']*1e-12/peak_tf:.1f}% of the {peak_tf} TFLOP/s DFMA probe, "
          f"{tot['bytes']/tot['t']*1e-9:.0f} GB/s = {100*tot['bytes']/tot['t']*1e-9/peak_gbs:.0f}% of the measured {peak_gbs} GB/s HBM copy peak")
    print(json.dumps({"flops_per_state": round(tot["flops"]/nstates, 1), "dram_bytes_per_state": round(tot["bytes"]/nstates, 1)}))


if __name__ == "__main__":
    main()
