#!/bin/bash
# A/B of the L2 evict-first policy on the image streams (SC_L2_STREAM_HINT): two bench runs each way + parity tests with it on.
# Needs a library built with SC_EXTRA_NVCC_FLAGS=-DSC_L2_STREAM_HINT_BUILD.  NOT yet run to completion: the one attempt in
# round 1 hit the end of the GPU budget (a cold `import torch` on a fresh box takes about a minute: keep the timeouts generous).
python -c "import torch" > /dev/null 2>&1
for h in 0 1 0 1; do
  SC_L2_STREAM_HINT=$h timeout 240 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab_hint_$h.json
  python - "$h" <<'PY'
import json, sys
h = sys.argv[1]
d = json.loads(open(f"gpurun_out/ab_hint_{h}.json").read().strip().splitlines()[-1])
print("hint", h, round(d["value"]), "samples/s", d["ms_per_step"], "ms/step; analysis launch", d["roofline"]["ms_per_launch"], "ms; sm", d["clocks"]["sm_mhz"])
PY
done
SC_L2_STREAM_HINT=1 timeout 120 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
