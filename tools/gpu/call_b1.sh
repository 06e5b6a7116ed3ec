#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_latency_check.py 1,2,4,32 2>&1 | grep "B="
python -m pytest tests/test_gpu_parity.py -x -q -k "small_batch or golden or odd or ffhq or parameter_gradients or graph or dlatent or wgrad or batch8 or bench_workload" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 240 --csv --log-file gpurun_out/lat_b1_fin.csv python tools/gpu_latency_check.py 1 > gpurun_out/lat_ncu.log 2>&1
