#!/bin/bash
# round 2, GPU call Z (final build: four resident CTAs for the sweep stages and the simple-pair narrow kernel):
# parity subset, full bench line + reference arm, other workloads, launch lists
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "xfrc or sensors2 or sensordata or golden_qfrc or golden_discrete or live_reference_4096 or overflow" > gpurun_out/z_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/z_summary.txt
python bench.py > gpurun_out/z_bench_headline.json 2> gpurun_out/z_bench_headline.err
python bench.py --impl reference > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err
python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline --no-other-configs > gpurun_out/z_bench_nocontact.json 2> gpurun_out/z_bench_nocontact.err
python bench.py --steps 10 --warmup 3 --workload humanoid_contact_elliptic --no-cpu-baseline --no-other-configs > gpurun_out/z_bench_elliptic.json 2> gpurun_out/z_bench_elliptic.err
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/z_bench_h22.json 2> gpurun_out/z_bench_h22.err
MJB_JIT_DEFINES="MJBS_SMOOTH_CTAS=5" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/z_bench_smooth5.json 2> gpurun_out/z_bench_smooth5.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,launch__registers_per_thread
ncu --metrics $M --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/z_launches_headline.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/z_ncu1.log 2>&1
ncu --metrics $M --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/z_launches_nocontact.csv python bench.py --steps 2 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline --no-other-configs > gpurun_out/z_ncu2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/z_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        if j.get("impl") == "reference": print(f, "reference value %.4g"%j["value"]); continue
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], j.get("kernel_mode","")[:34], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/z_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/z_tests.log | tail -n 8
