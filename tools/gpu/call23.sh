#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for lib in new pm1 pm2; do
  if [ $lib = new ]; then unset SGR_LIB; else export SGR_LIB=$PWD/tools/ab/libsgr_$lib.so; fi
  for f in 128 1; do
  SGR_FUSE_FIR=$f SGR_BENCH_CHILD=1 python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ab_${lib}_$rep.json 2>/dev/null
  python - $lib $rep $f <<'P'
import json, sys
d=json.loads(open('gpurun_out/ab_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
r=d['roofline']
print(sys.argv[1], 'FUSE', sys.argv[3], sys.argv[2], 'value %.0f ms %.3f sustained %.3f issued %.3f kernel_ms %.3f fir %.3f' % (d['value'], d['ms_per_step'], d['sustained']['ms_per_step'], r['issued_frac'], r['kernel_ms_per_step'], r['hbm_pass']['ms_per_step']), [l['ms'] for l in d['layers']][4:])
P
done; done; done
