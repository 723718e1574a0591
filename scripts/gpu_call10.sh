#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
bash scripts/gpu_ab.sh SC_ANA2 "0 1"
bash scripts/gpu_ab.sh SC_QUAD3 "0 1"
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-graph > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | tail -8
