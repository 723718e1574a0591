#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast.py -q -m gpu --timeout 600 2>&1 | tail -6
bash scripts/gpu_ab.sh SC_QM_TMA "0 1 0 1"
bash scripts/gpu_ab.sh SC_SYN_STORE_HINT "0 1"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ab_SC_QM_TMA_1.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("kernels"), indent=1))
print("analysis us", d["roofline"]["ms_per_launch"] * 1e3)
PY
