#!/bin/bash
mkdir -p gpurun_out
timeout 100 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on -k "regex:k_fused_analysis2|k_fused_synthesis|k_mode_gemm_quad3" -s 7 -c 3 -o gpurun_out/r02_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-graph > gpurun_out/ncu_full.log 2>&1
timeout 150 python bench.py --steps 50 --warmup 10 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 600 gpurun_out/r02_bench.json
