#!/bin/bash
# round-2 GPU call 4: MMA microbench (distinct operands / accumulator rotation), NT=128 interleave A/B, drop-in tests
mkdir -p gpurun_out
echo "== mma microbench 2"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/microbench/mma_rate.cu && /tmp/mma_rate 2>&1 | tee gpurun_out/mma_rate2.log
echo "== L10 interleave A/B"
python tools/gpu_layer_bench.py 32 "L10" 2>&1 | sed "s/^/new /"
SGR_DEBUG=128 python tools/gpu_layer_bench.py 32 "L10" 2>&1 | sed "s/^/old /"
python tools/gpu_layer_bench.py 32 "L12" 2>&1
echo "== dropin tests"; timeout 900 python -m pytest tests/test_reference_dropin.py tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -15
echo "== bench"; python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/bench_r2_c4.json 2>/dev/null; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2_c4.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'sustained', d['sustained']['ms_per_step'])
r=d['roofline']; print('roofline', {k:r[k] for k in ('achieved','frac','issued_frac','issued_frac_vs_burst','kernel_ms_per_step')})
print('layers', [(l['layer'],l['ms']) for l in d['layers']])
P
du -sh gpurun_out
