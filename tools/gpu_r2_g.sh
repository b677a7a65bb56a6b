#!/bin/bash
# round 2, GPU call G: compact scan tables + tree-level broadphase + per-model lists (config 5); all suites
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
python -m pytest tests -m gpu -q > gpurun_out/g_tests_generic.log 2>&1; echo "generic tests rc=$?" > gpurun_out/g_summary.txt
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline > gpurun_out/g_bench_h22.json 2> gpurun_out/g_bench_h22.err
MJB_CONTACT_PATH=pooled python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline > gpurun_out/g_bench_h22_pooled.json 2> gpurun_out/g_bench_h22_pooled.err
python tools/contact_queue_counters.py > gpurun_out/g_queue_counters.txt 2>&1
MJB_JIT=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_jit_contact.json 2> gpurun_out/g_bench_jit_contact.err
MJB_JIT=1 MJB_JIT_STAGE_COST=5000 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload humanoid_nocontact > gpurun_out/g_bench_jit_cost5000.json 2> gpurun_out/g_bench_jit_cost5000.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/g_summary.txt
grep -E "FAILED|passed|failed" gpurun_out/g_tests_generic.log | tail -n 12
tail -n 8 gpurun_out/g_queue_counters.txt
