#!/bin/bash
# round 2, GPU call E: no automatic contraction (exact kinematics / rows), explicit FMA in the backward sweeps
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_scale.py > gpurun_out/e_tests_generic.log 2>&1; echo "generic tests rc=$?" > gpurun_out/e_summary.txt
MJB_JIT=1 timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_scale.py > gpurun_out/e_tests_jit.log 2>&1; echo "jit tests rc=$?" >> gpurun_out/e_summary.txt
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q > gpurun_out/e_tests_scale.log 2>&1; echo "scale tests rc=$?" >> gpurun_out/e_summary.txt
for w in humanoid_nocontact humanoid_contact_pyramidal; do
  python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/e_bench_generic_$w.json 2> gpurun_out/e_bench_generic_$w.err
  MJB_JIT=1 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/e_bench_jit_$w.json 2> gpurun_out/e_bench_jit_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/e_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "launches", j["gpu_launches"], "e2e %.3g"%j["e2e"]["value"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/e_summary.txt
grep -E "FAILED|passed|failed" gpurun_out/e_tests_generic.log | tail -n 8
grep -E "FAILED|passed|failed" gpurun_out/e_tests_jit.log | tail -n 8
grep -E "FAILED|passed|failed" gpurun_out/e_tests_scale.log | tail -n 12
