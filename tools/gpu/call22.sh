#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for f in 128 0 256 512 1; do
  SGR_FUSE_FIR=$f SGR_BENCH_CHILD=1 python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/fuse_${f}_$rep.json 2>/dev/null
  python - $f $rep <<'P'
import json, sys
d=json.loads(open('gpurun_out/fuse_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
r=d['roofline']
print('FUSE', sys.argv[1], sys.argv[2], 'value %.0f ms %.3f sustained %.3f issued %.3f kernel_ms %.3f fir_ms %.3f' % (d['value'], d['ms_per_step'], d['sustained']['ms_per_step'], r['issued_frac'], r['kernel_ms_per_step'], r['hbm_pass']['ms_per_step']), [l['ms'] for l in d['layers']])
P
done; done
