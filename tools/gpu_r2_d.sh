#!/bin/bash
# round 2, GPU call D: wrench masks / nc loads / staged kernels; all test suites; launch list under ncu
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_scale.py > gpurun_out/d_tests_generic.log 2>&1; echo "generic tests rc=$?" > gpurun_out/d_summary.txt
MJB_JIT=1 timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_scale.py > gpurun_out/d_tests_jit.log 2>&1; echo "jit tests rc=$?" >> gpurun_out/d_summary.txt
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q > gpurun_out/d_tests_scale.log 2>&1; echo "scale tests rc=$?" >> gpurun_out/d_summary.txt
for w in humanoid_nocontact humanoid_contact_pyramidal; do
  python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/d_bench_generic_$w.json 2> gpurun_out/d_bench_generic_$w.err
  MJB_JIT=1 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/d_bench_jit_$w.json 2> gpurun_out/d_bench_jit_$w.err
  MJB_JIT=1 MJB_JIT_PHASES=smooth,scan,backward python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/d_bench_jitnoinertia_$w.json 2> gpurun_out/d_bench_jitnoinertia_$w.err
done
MJB_JIT=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_plain.log 2>&1 && \
MJB_JIT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/d_launches_jit.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/d_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "launches", j["gpu_launches"], "e2e %.3g"%j["e2e"]["value"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/d_summary.txt
tail -n 4 gpurun_out/d_tests_generic.log
tail -n 4 gpurun_out/d_tests_jit.log
grep -E "FAILED|passed|failed" gpurun_out/d_tests_scale.log | tail -n 12
