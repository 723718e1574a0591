#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 240 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -4
for n in 8 4; do
  SC_RESERVED_SMS=12 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python - "$n" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n{f}.json").read().strip().splitlines()[-1])
    print("N", f, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step graph", d["config"]["cuda_graph"], d["config"].get("cuda_graph_error"))
except Exception as e:
    print(f, "no result", e, open(f"gpurun_out/bench_n{f}.err").read()[-1500:])
PY
done
SC_ALLREDUCE=nccl timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n8_nccl.json").read().strip().splitlines()[-1])
    print("N 8 nccl", round(d["value"]), "samples/s", round(d["ms_per_step"], 5))
except Exception as e:
    print("nccl no result", e)
PY
