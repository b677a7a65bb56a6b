#!/bin/bash
# round 2, GPU call Y: final build check (xfrc fix, three resident CTAs for the sweep stages) + two more occupancy variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "xfrc or sensors2 or sensordata or golden_qfrc or golden_discrete or live_reference_4096" > gpurun_out/y_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/y_summary.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/y_bench_base.json 2> gpurun_out/y_bench_base.err
MJB_JIT_DEFINES="MJBS_SMOOTH_CTAS=4" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/y_bench_smooth4.json 2> gpurun_out/y_bench_smooth4.err
MJB_LIB=$PWD/mujoco_inversedynamicstest_b200/lib/variants/libmjb_narrow4.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/y_bench_narrow4.json 2> gpurun_out/y_bench_narrow4.err
python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline --no-other-configs > gpurun_out/y_bench_nocontact.json 2> gpurun_out/y_bench_nocontact.err
python bench.py --steps 10 --warmup 3 --workload humanoid_contact_elliptic --no-cpu-baseline --no-other-configs > gpurun_out/y_bench_elliptic.json 2> gpurun_out/y_bench_elliptic.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/y_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], j.get("kernel_mode","")[:40], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/y_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/y_tests.log | tail -n 8
