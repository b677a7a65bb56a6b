#!/bin/bash
# round 2, GPU call L: warp-per-state contact numbering (22 / 100 humanoids)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "humanoids22 or hundred or multi_tree or subwarp or overflow or edge_case" > gpurun_out/l_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/l_summary.txt
timeout 300 python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/l_bench_h22.json 2> gpurun_out/l_bench_h22.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/l_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/l_tests.log | tail -n 16
tail -n 3 gpurun_out/l_bench_h22.err
