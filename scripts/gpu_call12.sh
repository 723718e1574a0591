#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
rm -f gpurun_out/trace2.txt
SC_TRACE_FILE=gpurun_out/trace2.txt timeout 120 python scripts/trace_run.py
python scripts/show_trace.py gpurun_out/trace2.txt | cut -c1-260
