#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -4
for cfg in "p2p 12 12" "p2p 16 16" "p2p 8 8" "nccl 12 12"; do
  set -- $cfg
  SC_ALLREDUCE=$1 SC_ALLREDUCE_CTAS=$2 SC_RESERVED_SMS=$3 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n2_$1_$2.json 2> gpurun_out/bench_n2_$1_$2.err
  python - "$1_$2" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n2_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step graph", d["config"]["cuda_graph"], d["config"].get("cuda_graph_error"))
except Exception as e:
    print(f, "no result", e, open(f"gpurun_out/bench_n2_{f}.err").read()[-1500:])
PY
done
