#!/bin/bash
mkdir -p gpurun_out
echo "== layer bench (new issuer)"; python tools/gpu_layer_bench.py 32 2>&1
echo "== no loads"; SGR_DEBUG=20 python tools/gpu_layer_bench.py 32 2>&1 | grep -E "L6|L8|L10|L12"
echo "== failing test detail"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parameter_gradients_train_mode" 2>&1 | grep -E "Error|assert|^E " | head -20
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or parity or dlatent or independence or determinism" 2>&1 | tail -4
