#!/bin/bash
# quad2 bring-up: contraction tests, whole GPU tier, A/B of the two quad generations, launch list
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_contract.py -m gpu -q -x 2>&1 | tee gpurun_out/contract_tests.log | tail -25
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
for q in 1 2 1 2; do
  SC_QUAD=$q timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/ab_quad_$q.err | tail -1 > gpurun_out/ab_quad_$q.json
  python - "$q" <<'PY'
import json, sys
q = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_quad_{q}.json").read().strip().splitlines()[-1])
    print("quad", q, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step; graph", d["config"].get("cuda_graph"), "step frac", round(d["roofline"]["step"]["frac"], 4))
except Exception as e:   # noqa: BLE001
    print("quad", q, "no result:", e, open(f"gpurun_out/ab_quad_{q}.err").read()[-800:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | tail -25
