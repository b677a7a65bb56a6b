#!/bin/bash
# round 2, GPU call B: specialised kernels with per-body barriers; parity at scale
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
for w in humanoid_contact_pyramidal humanoid_nocontact; do
  MJB_JIT=1 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/b_bench_jit_$w.json 2> gpurun_out/b_bench_jit_$w.err
done
MJB_JIT=1 MJB_JIT_DEFINES="MJBS_INERTIA_THREADS=512,MJBS_BACKWARD_THREADS=256" python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/b_bench_jit_v2.json 2> gpurun_out/b_bench_jit_v2.err
MJB_JIT=1 MJB_JIT_DEFINES="MJB_BODY_SYNC()=" python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/b_bench_jit_nosync.json 2> gpurun_out/b_bench_jit_nosync.err
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q -x > gpurun_out/b_tests_scale.log 2>&1; echo "scale tests rc=$?" > gpurun_out/b_summary.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "e2e %.3g"%j["e2e"]["value"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 12 gpurun_out/b_tests_scale.log
