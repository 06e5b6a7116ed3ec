#!/bin/bash
# batch-1 latency A/B: scatter split-K cap (0 = off, 4, 8) with programmatic dependent launch in its default (auto) mode
mkdir -p gpurun_out
for cap in 0 4 8; do
  echo "SPLITK=$cap"
  SGR_UP_SPLITK=$cap python tools/gpu_latency_check.py 1,2,4 2>&1 | grep "graph"
done
python -m pytest tests/test_gpu_parity.py -x -q -k "small_batch or golden or odd or ffhq or parameter_gradients or graph or dlatent_vs" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 240 --csv --log-file gpurun_out/lat_b1_split8.csv python tools/gpu_latency_check.py 1 > gpurun_out/lat_ncu.log 2>&1
