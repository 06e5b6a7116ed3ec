#!/bin/bash
echo "== tests (gradients / caches / drop-in)"; timeout 1200 python -m pytest tests -m gpu -q -x -k "parameter or wgrad or optimize or weight_cache or graph or dlatent or train_step or reference or dA" 2>&1 | tail -4
echo "== optimize_g B=1"; python tools/gpu_optimize_g_bench.py 1 2>&1 | tail -1
echo "== optimize_g B=16"; python tools/gpu_optimize_g_bench.py 16 2>&1 | tail -1
python tools/gpu_optimize_g_cpuprofile.py 1 2>&1 | grep "^sync"
