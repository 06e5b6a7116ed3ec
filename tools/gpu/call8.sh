#!/bin/bash
# which load stream bounds the GEMM kernels?  debug 4: weight ring loaded once, 16: activation ring loaded once (timing only)
for d in 0 4 16 20; do SGR_DEBUG=$d python tools/gpu_layer_bench.py 32 2>&1 | sed "s/^/DEBUG=$d /"; done
