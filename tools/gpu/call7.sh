#!/bin/bash
mkdir -p gpurun_out
echo "== mma layout sweep"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/microbench/mma_rate.cu && /tmp/mma_rate l 2>&1 | tee gpurun_out/mma_layout.log
echo "== upfirdn2d kernels"
for m in 1 2; do SGR_UPFIRDN_ROWS=$m python tools/gpu_upfirdn_bench.py; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_dropin.py -m gpu -q -k "upfirdn or native_ops" 2>&1 | tail -3
