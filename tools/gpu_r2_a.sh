#!/bin/bash
# round 2, GPU call A: generic path after the loop refactor, then the model-specialised kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_tests_generic.log 2>&1; echo "generic tests rc=$?" > gpurun_out/a_summary.txt
MJB_JIT=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/a_tests_jit.log 2>&1; echo "jit tests rc=$?" >> gpurun_out/a_summary.txt
for w in humanoid_contact_pyramidal humanoid_nocontact; do
  python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/a_bench_generic_$w.json 2> gpurun_out/a_bench_generic_$w.err
  MJB_JIT=1 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/a_bench_jit_$w.json 2> gpurun_out/a_bench_jit_$w.err
done
tail -3 gpurun_out/a_tests_generic.log gpurun_out/a_tests_jit.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/a_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "e2e %.3g"%j["e2e"]["value"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
