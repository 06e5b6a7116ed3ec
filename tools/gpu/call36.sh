#!/bin/bash
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
for lin in 1 0 1 0; do
SGR_UP_LINEAR=$lin python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > /tmp/b_$lin.json 2>/dev/null
python - <<P
import json
d=json.loads(open('/tmp/b_$lin.json').read().strip().splitlines()[-1])
print('linear $lin: ms %.3f value %.0f issued %.3f' % (d['ms_per_step'], d['value'], d['roofline']['issued_frac']), ' '.join('%s:%.3f' % (l.get('layer', i), l['ms']) for i, l in enumerate(d['layers'])))
P
done
