#!/bin/bash
timeout 60 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tucker or tfno" 2>&1 | tail -3
