#!/bin/bash
# per-kernel launch lists for lib/libmjb.so and every lib/variants/libmjb_*.so
run() {
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -c 40 --csv --log-file gpurun_out/lv_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lv_$1.log 2>&1
}
run base
for f in mujoco_inversedynamicstest_b200/lib/variants/libmjb_*.so; do
  n=$(basename $f .so); n=${n#libmjb_}
  MJB_LIB=$PWD/$f run $n
done
