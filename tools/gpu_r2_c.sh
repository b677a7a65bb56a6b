#!/bin/bash
# round 2, GPU call C: staged specialised kernels (stage-size sweep), full GPU suite on them, scale parity
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
for w in humanoid_nocontact humanoid_contact_pyramidal; do
  MJB_JIT=1 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/c_bench_jit_$w.json 2> gpurun_out/c_bench_jit_$w.err
done
for cost in 2300 7000 14000; do
  MJB_JIT=1 MJB_JIT_STAGE_COST=$cost python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/c_bench_jit_cost$cost.json 2> gpurun_out/c_bench_jit_cost$cost.err
done
MJB_JIT=1 MJB_JIT_DEFINES="MJBS_SMOOTH_CTAS=3,MJBS_INERTIA_CTAS=4" python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/c_bench_jit_occ.json 2> gpurun_out/c_bench_jit_occ.err
MJB_JIT=1 timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_scale.py > gpurun_out/c_tests_jit.log 2>&1; echo "jit tests rc=$?" > gpurun_out/c_summary.txt
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q > gpurun_out/c_tests_scale.log 2>&1; echo "scale tests rc=$?" >> gpurun_out/c_summary.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/c_summary.txt
tail -n 5 gpurun_out/c_tests_jit.log
grep -E "FAILED|passed|failed" gpurun_out/c_tests_scale.log | tail -n 30
