#!/bin/bash
# usage: scripts/run_scaling.sh N   -> gpurun_out/bench_{C4,C5,C2}_n$N.json  (run under gpurun --gpus N)
N=$1
for c in C4 C5 C2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --config $c --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_${c}_n${N}_err.log | tail -1 > gpurun_out/bench_${c}_n${N}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${c}_n${N}.json"))
    print("${c} N=${N}:", d["metric"], d["value"], "ms/step", d["ms_per_step"], "collective", json.dumps(d.get("collective"))[:300], "train", json.dumps(d.get("train", {}).get("collective"))[:300], d.get("train", {}).get("ms_per_step"))
except Exception as e:
    print("${c} N=${N}: no json", e)
    print(open("gpurun_out/bench_${c}_n${N}_err.log").read()[-1500:])
PY
done
