#!/bin/bash
# round 2, GPU call F: fused subtree stages (phase_tree)
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
MJB_JIT=1 timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_scale.py > gpurun_out/f_tests_jit.log 2>&1; echo "jit tests rc=$?" > gpurun_out/f_summary.txt
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q > gpurun_out/f_tests_scale.log 2>&1; echo "scale tests rc=$?" >> gpurun_out/f_summary.txt
for w in humanoid_nocontact humanoid_contact_pyramidal; do
  MJB_JIT=1 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/f_bench_jit_$w.json 2> gpurun_out/f_bench_jit_$w.err
  MJB_JIT=1 MJB_JIT_PHASES=smooth,inertia,scan,backward python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/f_bench_jitnotree_$w.json 2> gpurun_out/f_bench_jitnotree_$w.err
done
MJB_JIT=1 MJB_JIT_FUSED_COST=12000 python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/f_bench_jit_fused12k.json 2> gpurun_out/f_bench_jit_fused12k.err
MJB_JIT=1 MJB_JIT_DEFINES="MJBS_TREE_CTAS=3" python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/f_bench_jit_ctas3.json 2> gpurun_out/f_bench_jit_ctas3.err
MJB_JIT=1 MJB_JIT_DEFINES="MJBS_TREE_CTAS=1" python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline > gpurun_out/f_bench_jit_ctas1.json 2> gpurun_out/f_bench_jit_ctas1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "launches", j["gpu_launches"], "e2e %.3g"%j["e2e"]["value"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/f_summary.txt
grep -E "FAILED|passed|failed" gpurun_out/f_tests_jit.log | tail -n 12
grep -E "FAILED|passed|failed" gpurun_out/f_tests_scale.log | tail -n 12
