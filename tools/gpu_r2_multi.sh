#!/bin/bash
# round 2, multi-GPU call: N = $1 ranks. Headline bench (value + e2e), copy ceiling probe, config 5 strong-scaling point
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then RUN="python"; PORTARG=""; fi
port=29611
launch() { if [ "$N" = "1" ]; then python "$@"; else $RUN --master-port $port "$@"; port=$((port+1)); fi; }
launch tools/pcie_probe.py > gpurun_out/m${N}_pcie_probe.json 2> gpurun_out/m${N}_pcie_probe.err
launch bench.py --gpus $N --steps 10 --warmup 3 --no-other-configs > gpurun_out/m${N}_bench_headline.json 2> gpurun_out/m${N}_bench_headline.err
# config 5 as BASELINE states it: 2^24 humanoid instances = 762,600 scene states of 22 humanoids in total, sharded over N GPUs
B=$((762600 / N))
launch bench.py --gpus $N --steps 3 --warmup 3 --workload humanoids22 --batch $B --no-cpu-baseline --no-other-configs > gpurun_out/m${N}_bench_h22.json 2> gpurun_out/m${N}_bench_h22.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/m${N}_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        if "probe" in j: print(f, {k:(round(v,2) if isinstance(v,float) else v) for k,v in j.items() if k!="probe"})
        else: print(f, "ms/step %.3f"%j["ms_per_step"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], j["config"].get("states_per_gpu"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 gpurun_out/m${N}_*.err
