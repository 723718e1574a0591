#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -15
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k transform 2>&1 | tail -3
for mode in p2p nccl; do
  SC_ALLREDUCE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err
  python - "$mode" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n2_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step graph", d["config"]["cuda_graph"], d["config"].get("cuda_graph_error"))
except Exception as e:
    print(f, "no result", e, open(f"gpurun_out/bench_n2_{f}.err").read()[-1500:])
PY
done
