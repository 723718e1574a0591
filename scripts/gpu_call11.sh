#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "transform or golden" 2>&1 | tail -5
# 2-GPU: in-graph overlapped all-reduce
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
SC_RESERVED_SMS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > gpurun_out/bench_n2_r0.json 2> gpurun_out/bench_n2_r0.err
python - <<'PY'
import json
for f in ("bench_n2", "bench_n2_r0"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step graph", d["config"]["cuda_graph"], d["config"].get("cuda_graph_error"), "reserved", d["config"]["reserved_sms"])
    except Exception as e:
        print(f, "no result", e)
PY
rm -f gpurun_out/trace2.txt
SC_TRACE_FILE=gpurun_out/trace2.txt timeout 120 python scripts/trace_run.py
python scripts/show_trace.py gpurun_out/trace2.txt | cut -c1-260
