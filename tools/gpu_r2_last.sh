#!/bin/bash
# per-state eq_active and adhesion transmission on the GPU, headline sanity
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q -k "eq_active or adhesion or xfrc" > gpurun_out/last_tests.log 2>&1; echo "tests rc=$?"
grep -E "FAILED|passed|failed|Error" gpurun_out/last_tests.log | tail -n 8
timeout 50 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
python -c "
import json; j=json.loads(open('gpurun_out/last_bench.json').read().strip().splitlines()[-1]); print('ms/step %.3f'%j['ms_per_step'], j['kernel_mode'], {k['kernel']: round(k['ms_per_step'],3) for k in j['kernels']})"
