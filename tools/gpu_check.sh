#!/bin/bash
# run on the GPU box: parity tests, a bench line, and the ncu launch list of one bench step
tag=${1:-run}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t_$tag.log
python bench.py --no-cpu-baseline > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 40 --csv --log-file gpurun_out/l_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
echo "=== GPU TESTS: $(tail -1 gpurun_out/t_$tag.log)"; cut -c1-200 gpurun_out/b_$tag.json
