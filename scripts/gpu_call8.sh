#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 120 python scripts/probe_hbm.py 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_contract.py -q -m gpu --timeout 300 -x 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast.py -q -m gpu --timeout 600 2>&1 | tail -12
bash scripts/gpu_ab.sh SC_ANA2 "0 1 0 1"
SC_ANA2=0 bash scripts/gpu_ab.sh SC_QUAD3 "0 1"
SC_ANA2=0 bash scripts/gpu_ab.sh SC_WIDE_BOXES "0 1"
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-graph > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | tail -8
SC_EXTRA_NVCC_FLAGS=-DSC_TRACE_QUAD python -m neuraloperator_b200.build --force > /dev/null 2>&1
rm -f gpurun_out/q3trace.txt
SC_TRACE_FILE=gpurun_out/q3trace.txt timeout 120 python scripts/trace_run.py
