#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mode_gemm_quad2 -s 6 -c 3 -o gpurun_out/r02_quad2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_quad2.log 2>&1
tail -3 gpurun_out/ncu_quad2.log
ls -la gpurun_out/*.ncu-rep | tail -3
