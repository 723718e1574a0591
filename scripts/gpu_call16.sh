#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -4
for cfg in "1 12" "1 8" "1 4" "0 12"; do
  set -- $cfg
  SC_ALLREDUCE_TMA=$1 SC_ALLREDUCE_CTAS=$2 SC_RESERVED_SMS=12 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n2_tma$1_$2.json 2> gpurun_out/bench_n2_tma$1_$2.err
  python - "tma$1_$2" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n2_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step graph", d["config"]["cuda_graph"], d["config"].get("cuda_graph_error"))
except Exception as e:
    print(f, "no result", e, open(f"gpurun_out/bench_n2_{f}.err").read()[-1500:])
PY
done
