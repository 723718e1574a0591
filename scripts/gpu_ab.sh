#!/bin/bash
# usage: gpu_ab.sh ENVVAR "v1 v2 v1 v2"   -- bench.py A/B over the values of one environment switch
var=$1; vals=$2
for q in $vals; do
  env $var=$q timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-configs 2> gpurun_out/ab_${var}_$q.err | tail -1 > gpurun_out/ab_${var}_$q.json
  python - "$var" "$q" <<'PY'
import json, sys
var, q = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/ab_{var}_{q}.json").read().strip().splitlines()[-1])
    print(var, q, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step; graph", d["config"].get("cuda_graph"), "step frac", round(d["roofline"]["step"]["frac"], 4))
except Exception as e:   # noqa: BLE001
    print(var, q, "no result:", e, open(f"gpurun_out/ab_{var}_{q}.err").read()[-800:])
PY
done
