#!/bin/bash
# round 2, last GPU call: the whole GPU suite + smoke on the final build
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.jsonl
timeout 340 python -m pytest tests -m gpu -q -x > gpurun_out/final_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/final_summary.txt
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_summary.txt
cat gpurun_out/final_summary.txt; tail -1 gpurun_out/final_smoke.log
grep -E "FAILED|passed|failed|Error" gpurun_out/final_tests.log | tail -n 8
