#!/bin/bash
# round 2, GPU call T: 128-byte aligned hot tables in the model blob; new output tests, bench lines, reference arm,
# ncu launch lists (headline / contact-free / 22 humanoids), ncu --set full of one headline step exported as CSV
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
python -m pytest tests -m gpu -q -k "camlight or transmission or sensordata or sensors2 or outputs_only or energy" > gpurun_out/t_tests_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/t_summary.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/t_bench_headline10.json 2> gpurun_out/t_bench_headline10.err
python bench.py > gpurun_out/t_bench_headline.json 2> gpurun_out/t_bench_headline.err
python bench.py --impl reference > gpurun_out/t_bench_reference.json 2> gpurun_out/t_bench_reference.err
python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline --no-other-configs > gpurun_out/t_bench_nocontact.json 2> gpurun_out/t_bench_nocontact.err
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/t_bench_h22.json 2> gpurun_out/t_bench_h22.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,launch__registers_per_thread
ncu --metrics $M --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/t_launches_headline.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/t_ncu1.log 2>&1
ncu --metrics $M --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/t_launches_nocontact.csv python bench.py --steps 2 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline --no-other-configs > gpurun_out/t_ncu2.log 2>&1
ncu --metrics $M --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/t_launches_h22.csv python bench.py --steps 1 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/t_ncu3.log 2>&1
timeout 420 ncu --set full --clock-control none -s 48 -c 16 -o /tmp/t_full_headline -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/t_ncu4.log 2>&1
ncu -i /tmp/t_full_headline.ncu-rep --page raw --csv > gpurun_out/t_full_headline_raw.csv 2> gpurun_out/t_ncu5.log
ls -la /tmp/t_full_headline.ncu-rep >> gpurun_out/t_summary.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/t_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        if j.get("impl") == "reference": print(f, "reference value %.4g"%j["value"], j.get("cpu_baseline")); continue
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/t_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/t_tests_new.log | tail -n 8
du -sh gpurun_out
