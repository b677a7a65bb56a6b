#!/bin/bash
# round 2, GPU call U: cold tables behind the staged part of the model blob (phase kernels' occupancy), implicit mj_discreteAcc
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "camlight or transmission or sensordata or sensors2 or outputs_only or energy or implicit or invdiscrete" > gpurun_out/u_tests_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/u_summary.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/u_bench_headline10.json 2> gpurun_out/u_bench_headline10.err
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/u_bench_h22.json 2> gpurun_out/u_bench_h22.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/u_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/u_tests_new.log | tail -n 8
