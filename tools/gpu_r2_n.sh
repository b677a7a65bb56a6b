#!/bin/bash
# round 2, GPU call N: energy outputs, FD sensor Jacobians, scan hot loop v4; all suites
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
python -m pytest tests -m gpu -q > gpurun_out/n_tests_generic.log 2>&1; echo "generic tests rc=$?" > gpurun_out/n_summary.txt
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/n_bench_h22.json 2> gpurun_out/n_bench_h22.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/n_bench_headline.json 2> gpurun_out/n_bench_headline.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,launch__registers_per_thread
ncu --metrics $M --clock-control none -s 50 -c 36 --csv --log-file gpurun_out/n_launches_h22.csv python bench.py --steps 2 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/n_ncu4.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/n_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/n_tests_generic.log | tail -n 16
