#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_fast.py -q -m gpu --timeout 600 -x 2>&1 | tee gpurun_out/fast_tests.log | tail -40
