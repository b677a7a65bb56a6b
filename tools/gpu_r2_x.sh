#!/bin/bash
# round 2, GPU call X: resident-CTA variants of the specialised sweep kernels (MJB_JIT_DEFINES), new xfrc / rangefinder / tendon-sensor tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "xfrc or sensors2 or sensordata" > gpurun_out/x_tests_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/x_summary.txt
i=0
for D in "" "MJBS_SMOOTH_CTAS=3" "MJBS_INERTIA_CTAS=4" "MJBS_INERTIA_CTAS=2" "MJBS_SMOOTH_CTAS=3,MJBS_INERTIA_CTAS=4" "MJBS_BACKWARD_CTAS=6" "MJBS_SMOOTH_CTAS=1"; do
  MJB_JIT_DEFINES="$D" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/x_bench_$i.json 2> gpurun_out/x_bench_$i.err
  echo "$i: $D" >> gpurun_out/x_summary.txt
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/x_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], j.get("kernel_mode","")[:40], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/x_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/x_tests_new.log | tail -n 8
