#!/bin/bash
# round 2, GPU call O: pair-organised scan, narrow grid for scenes, inertia prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "humanoids22 or hundred or multi_tree or overflow or edge_case or inertia or fd_sensor or 1m_states" > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/o_summary.txt
python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/o_bench_h22.json 2> gpurun_out/o_bench_h22.err
MJB_SCAN=list python bench.py --steps 5 --warmup 3 --workload humanoids22 --no-cpu-baseline --no-other-configs > gpurun_out/o_bench_h22_list.json 2> gpurun_out/o_bench_h22_list.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/o_bench_headline.json 2> gpurun_out/o_bench_headline.err
MJB_JIT_DEFINES="MJB_NO_INERTIA_PREFETCH=1" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/o_bench_headline_noprefetch.json 2> gpurun_out/o_bench_headline_noprefetch.err
python bench.py --steps 10 --warmup 3 --workload humanoid_nocontact --no-cpu-baseline --no-other-configs > gpurun_out/o_bench_nocontact.json 2> gpurun_out/o_bench_nocontact.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/o_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "launches", j["gpu_launches"], j.get("kernel_mode","")[:30], {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/o_summary.txt
grep -E "FAILED|passed|failed|Error" gpurun_out/o_tests.log | tail -n 16
