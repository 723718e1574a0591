#!/bin/bash
timeout 75 python - <<'PY'
import subprocess, sys, time
t0 = time.time()
import __graft_entry__ as g
g.smoke()
print("smoke seconds", round(time.time() - t0, 1), flush=True)
import pytest
sys.exit(pytest.main(["tests/test_gpu_parity.py", "-q", "-m", "gpu", "-x", "-k", "tucker or tfno", "-p", "no:cacheprovider"]))
PY
