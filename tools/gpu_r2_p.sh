#!/bin/bash
# round 2, GPU call P: compact survivor masks for the warp-per-state scans
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "humanoids22 or hundred or multi_tree or overflow or edge_case" > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/p_summary.txt
MJB_CONTACT_PATH=pooled timeout 600 python -m pytest tests -m gpu -q -k "humanoids22 and (discrete or qfrc)" > gpurun_out/p_tests_pooled.log 2>&1; echo "pooled tests rc=$?" >> gpurun_out/p_summary.txt
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/p_bench_h22.json 2> gpurun_out/p_bench_h22.err
MJB_SCAN=list python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/p_bench_h22_list.json 2> gpurun_out/p_bench_h22_list.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/p_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/p_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/p_tests.log gpurun_out/p_tests_pooled.log | tail -n 16
