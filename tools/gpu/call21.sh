#!/bin/bash
mkdir -p gpurun_out
echo "== layer bench"; python tools/gpu_layer_bench.py 32 2>&1
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
for rep in 1 2; do for lib in prev new; do
  if [ $lib = new ]; then unset SGR_LIB; else export SGR_LIB=$PWD/tools/ab/libsgr_$lib.so; fi
  python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ab_${lib}_$rep.json 2>/dev/null
  python - $lib $rep <<'P'
import json, sys
d=json.loads(open('gpurun_out/ab_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
r=d['roofline']
print(sys.argv[1], sys.argv[2], 'value %.0f ms %.3f sustained %.3f issued %.3f kernel_ms %.3f fir %.3f' % (d['value'], d['ms_per_step'], d['sustained']['ms_per_step'], r['issued_frac'], r['kernel_ms_per_step'], r['hbm_pass']['ms_per_step']), [l['ms'] for l in d['layers']])
P
done; done
