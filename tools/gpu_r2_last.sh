#!/bin/bash
# adhesion transmission on the GPU
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -k "adhesion or transmission or fd" > gpurun_out/last_tests.log 2>&1; echo "tests rc=$?"
grep -E "FAILED|passed|failed|Error" gpurun_out/last_tests.log | tail -n 8
