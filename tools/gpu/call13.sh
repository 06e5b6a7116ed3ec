#!/bin/bash
mkdir -p gpurun_out
echo "== layer bench (issuers v2: halo + scatter + modconv)"; python tools/gpu_layer_bench.py 32 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_r2_c13.log
echo "== bench"; python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 > gpurun_out/bench_r2_c13.json 2>/dev/null; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2_c13.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.3f issued_frac %.3f (burst %.3f) kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['ms_per_step'], r['issued_frac'], r['issued_frac_vs_burst'], r['kernel_ms_per_step']), [l['ms'] for l in d['layers']])
print('train', json.dumps(d['train_step'])[:500])
print('strong', d['strong_scaling']['value'])
P
