"""Refresh profiles/flops_per_state.json from an ncu launch list (tools/gpu_check.sh):
    python tools/update_counts.py gpurun_out/l_TAG.csv [workload]"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "humanoid_contact_pyramidal"
    nst = float(sys.argv[3]) if len(sys.argv) > 3 else 1048576.0
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    d = OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    # every distinct kernel launches once per step (one chunk): per step and group, the sum over the
    # group's kernels of their per-launch averages over the captured launches
    per = OrderedDict()
    for (_, k), m in d.items():
        e = per.setdefault(k, {"n": 0, "flops": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0})
        e["n"] += 1
        e["flops"] += (m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] +
                       m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] +
                       2 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]) / nst
        e["dram_bytes"] += (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / nst
        e["warp_inst"] += m.get("smsp__inst_executed.sum", 0.0) / nst
    K = {}
    for k, pe in per.items():
        name = k.split("(")[0].replace("void ", "").split("<")[0].replace("_kernel", "").replace("mjb::", "")
        if name.startswith("mjbs_"):          # model-specialised stages: mjbs_smooth_3 -> smooth
            name = name[5:].rstrip("0123456789").rstrip("_")
        if "soa" in name or "aos" in name or "probe" in name or "count_nonzero" in name:
            continue                           # boundary transposes / probes are not part of the step
        # the contact phase = narrow + index + rows (+ the pooled fallback kernel)
        if name.startswith("contact") and name != "contact_scan":
            name = "contact"
        e = K.setdefault(name, {"flops": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0})
        for key in e:
            e[key] += pe[key] / pe["n"]
    for e in K.values():
        for k in e:
            e[k] = round(e[k], 1)
    out = os.path.join(ROOT, "profiles", "flops_per_state.json")
    rec = json.load(open(out)) if os.path.exists(out) else {}
    rec[workload] = {
        "flops_per_state": round(sum(e["flops"] for e in K.values()), 1),
        "dram_bytes_per_state": round(sum(e["dram_bytes"] for e in K.values()), 1),
        "kernels": K,
        "source": "ncu launch list of one bench step (%s): per kernel, (dadd + dmul + 2*dfma thread "
                  "instructions) and dram__bytes_read+write, divided by the %d states of the step; "
                  "qM/qLD/qLDiagInv outputs on" % (os.path.basename(path), int(nst))}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec[workload]["kernels"]))


if __name__ == "__main__":
    main()
