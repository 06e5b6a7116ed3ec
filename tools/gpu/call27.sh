#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
for rep in 1 2; do for ts in 0 1; do
  SGR_TAIL_STREAM=$ts SGR_BENCH_CHILD=1 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ts_${ts}_$rep.json 2>/dev/null
  python - $ts $rep <<'P'
import json, sys
d=json.loads(open('gpurun_out/ts_%s_%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
r=d['roofline']
print('TAILSTREAM', sys.argv[1], sys.argv[2], 'value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.3f issued %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['ms_per_step'], r['issued_frac'], r['kernel_ms_per_step']))
P
done; done
python tools/gpu_latency_check.py 2>&1 | grep "B= 1"
SGR_TAIL_STREAM=0 python tools/gpu_latency_check.py 2>&1 | grep "B= 1"
