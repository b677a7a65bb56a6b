#!/bin/bash
# round 2, GPU call Z2: wrap-geom frames of output-only tendons kept only in runs that need them (arm26), parity subset
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "transmission or sensors2 or golden_qfrc or golden_discrete or tendon" > gpurun_out/z2_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/z2_summary.txt
python bench.py --steps 10 --warmup 3 --workload arm26 --no-cpu-baseline --no-other-configs > gpurun_out/z2_bench_arm26.json 2> gpurun_out/z2_bench_arm26.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/z2_bench_headline.json 2> gpurun_out/z2_bench_headline.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/z2_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], j.get("kernel_mode","")[:34], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/z2_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/z2_tests.log | tail -n 8
