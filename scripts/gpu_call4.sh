#!/bin/bash
# quad-major layout: whole GPU tier, A/B against the standard layout, timeline of the three contraction launches
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
for q in 0 1 0 1; do
  SC_QUAD_MAJOR=$q timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/ab_qm_$q.err | tail -1 > gpurun_out/ab_qm_$q.json
  python - "$q" <<'PY'
import json, sys
q = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_qm_{q}.json").read().strip().splitlines()[-1])
    print("quad-major", q, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step; graph", d["config"].get("cuda_graph"), "step frac", round(d["roofline"]["step"]["frac"], 4))
except Exception as e:   # noqa: BLE001
    print("quad-major", q, "no result:", e, open(f"gpurun_out/ab_qm_{q}.err").read()[-800:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | tail -8
# timeline build (scratch: the snapshot's library is rebuilt on the box only)
SC_EXTRA_NVCC_FLAGS=-DSC_TRACE_QUAD python -m neuraloperator_b200.build --force > /dev/null 2>&1
rm -f gpurun_out/q2trace.txt gpurun_out/q2trace_std.txt
SC_TRACE_FILE=gpurun_out/q2trace.txt timeout 120 python scripts/trace_run.py
SC_QUAD_MAJOR=0 SC_TRACE_FILE=gpurun_out/q2trace_std.txt timeout 120 python scripts/trace_run.py
grep -c quad2 gpurun_out/q2trace.txt
