#!/bin/bash
# One gpurun call: parity tests, smoke, bench, reference-on-GPU denominator, ncu launch list. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 ${PYTEST_ARGS:--x} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -${PYTEST_TAIL:-15} gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ -z "$SKIP_TORCH" ]; then timeout 600 python scripts/bench_torch_gpu.py --configs ${TORCH_CONFIGS:-1,2,4,5a} > gpurun_out/torch_gpu.log 2>&1; tail -5 gpurun_out/torch_gpu.log | cut -c1-400; fi
if [ -z "$SKIP_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | tail -25
fi
