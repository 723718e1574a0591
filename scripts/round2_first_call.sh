#!/bin/bash
# First GPU call of round 2.  BEFORE calling gpurun, build the experimental library locally so
# that it travels with the snapshot (no nvcc time on the GPU box):
#     SC_EXTRA_NVCC_FLAGS="-DSC_ROWS_KERNELS -DSC_L2_STREAM_HINT_BUILD" python -m neuraloperator_b200.build --force
# then:  gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# and afterwards rebuild the default library (python -m neuraloperator_b200.build --force) unless the experiments are adopted.
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1          # page the image in once (a cold import takes about a minute)
python -c "from neuraloperator_b200 import _lib; print(_lib.load().sc_build_info().decode())"
# 1. the never-run last-dim tensor-core kernels: every wait in them is bounded (2 s trap), but keep an outer timeout anyway
timeout 300 python -m pytest tests/test_gpu_rows.py -m gpu -x -q 2>&1 | tail -15
# 2. nothing else regressed with the experimental build
timeout 300 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
# 3. cfg-5a with and without the rows kernels (SC_ROWS=0 switches them off at run time)
for r in 0 1; do
  SC_ROWS=$r timeout 300 python scripts/bench_torch_gpu.py --configs 5a 2>&1 | grep '^5a' | cut -c1-260 | sed "s/^/SC_ROWS=$r /"
done
# 4. L2 evict-first policy on the image streams, headline config
for h in 0 1 0 1; do
  SC_L2_STREAM_HINT=$h timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/ab_hint_$h.err | tail -1 > gpurun_out/ab_hint_$h.json
  python - "$h" <<'PY'
import json, sys
h = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_hint_{h}.json").read().strip().splitlines()[-1])
    print("hint", h, round(d["value"]), "samples/s", round(d["ms_per_step"], 5), "ms/step; graph", d["config"].get("cuda_graph"))
except Exception as e:   # noqa: BLE001
    print("hint", h, "no result:", e, open(f"gpurun_out/ab_hint_{h}.err").read()[-600:])
PY
done
