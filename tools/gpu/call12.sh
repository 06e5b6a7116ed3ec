#!/bin/bash
echo "== layer bench (issuer v2)"; python tools/gpu_layer_bench.py 32 2>&1 | grep -E "L6|L8|L10|L12"
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
