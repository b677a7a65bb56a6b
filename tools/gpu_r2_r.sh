#!/bin/bash
# round 2, GPU call R: geom position / z-axis vectors; all suites
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
python -m pytest tests -m gpu -q > gpurun_out/r_tests_generic.log 2>&1; echo "generic tests rc=$?" > gpurun_out/r_summary.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r_bench_headline.json 2> gpurun_out/r_bench_headline.err
python bench.py --steps 10 --warmup 3 --workload humanoid_contact_elliptic --no-cpu-baseline --no-other-configs > gpurun_out/r_bench_elliptic.json 2> gpurun_out/r_bench_elliptic.err
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/r_bench_h22.json 2> gpurun_out/r_bench_h22.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/r_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/r_tests_generic.log | tail -n 16
