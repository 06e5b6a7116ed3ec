#!/bin/bash
mkdir -p gpurun_out
echo "== parity (PDL on)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
for rep in 1 2; do for pdl in 0 1; do
  SGR_PDL=$pdl SGR_BENCH_CHILD=1 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/pdl_${pdl}_$rep.json 2>/dev/null
  python - $pdl $rep <<'P'
import json, sys
d=json.loads(open('gpurun_out/pdl_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
r=d['roofline']
print('PDL', sys.argv[1], sys.argv[2], 'value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.3f issued %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['ms_per_step'], r['issued_frac'], r['kernel_ms_per_step']))
P
done; done
