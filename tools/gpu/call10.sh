#!/bin/bash
mkdir -p gpurun_out
echo "== mma power test"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/microbench/mma_rate.cu && /tmp/mma_rate w 2>&1 | tee gpurun_out/mma_power.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r2_c10.log
echo "== bench (A/B NT=128 order on the same box)"
for d in 0 128; do SGR_DEBUG=$d python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/bench_r2_c10_$d.json 2>/dev/null; python - $d <<'P'
import json, sys
d=json.loads(open('gpurun_out/bench_r2_c10_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']
print('DEBUG', sys.argv[1], 'value %.0f ms %.3f sustained %.3f issued_frac %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['sustained']['ms_per_step'], r['issued_frac'], r['kernel_ms_per_step']), [l['ms'] for l in d['layers']])
P
done
