#!/bin/bash
for d in 20 276 1044 1300; do SGR_DEBUG=$d python tools/gpu_layer_bench.py 32 "L8" 2>&1 | sed "s/^/DEBUG=$d /"; SGR_DEBUG=$d python tools/gpu_layer_bench.py 32 "L10" 2>&1 | sed "s/^/DEBUG=$d /"; done
