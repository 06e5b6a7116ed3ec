#!/bin/bash
for f in 1 0; do echo "== parity with SGR_FUSE_FIR=$f"; SGR_FUSE_FIR=$f timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2; done
echo "== parity with SGR_PDL=1"; SGR_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or parity or determinism or graph" 2>&1 | tail -2
echo "== sweep 256 cm2"; python tools/gpu_sweep.py 256 2 32 2>&1 | tail -2
