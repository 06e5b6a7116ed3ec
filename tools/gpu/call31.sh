#!/bin/bash
mkdir -p gpurun_out
echo "== optimize_g B=1"; python tools/gpu_optimize_g_bench.py 1 2>&1 | tail -1
echo "== optimize_g B=16"; python tools/gpu_optimize_g_bench.py 16 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/optg_b1.csv python tools/gpu_optimize_g_bench.py 1 > /dev/null 2>&1; wc -l gpurun_out/optg_b1.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/optg_b16.csv python tools/gpu_optimize_g_bench.py 16 > /dev/null 2>&1; wc -l gpurun_out/optg_b16.csv
