#!/bin/bash
# usage: scripts/run_scaling3.sh N [configs...]  -> gpurun_out/bench3_{C4,C2,...}_n$N.json  (run under gpurun --gpus N)
N=$1; shift
CFGS=${@:-C4 C2}
for c in $CFGS; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --config $c --steps 10 --warmup 3 2>gpurun_out/bench3_${c}_n${N}_err.log | tail -1 > gpurun_out/bench3_${c}_n${N}.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --config $c --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench3_${c}_n${N}_err.log | tail -1 > gpurun_out/bench3_${c}_n${N}.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench3_${c}_n${N}.json"))
    t = d.get("train", {})
    print("${c} N=${N}:", d["metric"], d["value"], "ms/step", d["ms_per_step"], "eager", d.get("eager_ms_per_step"), "collective", json.dumps(d.get("collective"))[:300],
          "| train ms", t.get("ms_per_step"), "eager", t.get("eager_ms_per_step"), json.dumps(t.get("collective"))[:300])
except Exception as e:
    print("${c} N=${N}: no json", e)
    print(open("gpurun_out/bench3_${c}_n${N}_err.log").read()[-2500:])
PY
done
