#!/bin/bash
# round 2, GPU call V: occupancy variants of the contact numbering / rows / narrow kernels (model from global memory)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/v_bench_base.json 2> gpurun_out/v_bench_base.err
for f in mujoco_inversedynamicstest_b200/lib/variants/libmjb_*.so; do
  n=$(basename $f .so); n=${n#libmjb_}
  MJB_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/v_bench_$n.json 2> gpurun_out/v_bench_$n.err
done
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum
for n in idx_rows5g idx_narrowg; do
MJB_LIB=$PWD/mujoco_inversedynamicstest_b200/lib/variants/libmjb_$n.so ncu --metrics $M --clock-control none -k regex:contact -s 12 -c 16 --csv --log-file gpurun_out/v_launches_$n.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/v_ncu_$n.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f"%j["ms_per_step"], "parity", j.get("parity"), {k["kernel"]: round(k["ms_per_step"],3) for k in j["kernels"]})
    except Exception as e:
        print(f, "ERR", e)
PY
