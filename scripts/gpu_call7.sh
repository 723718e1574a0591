#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 120 python scripts/probe_tma.py 2>&1 | tail -8
