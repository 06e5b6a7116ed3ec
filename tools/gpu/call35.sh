#!/bin/bash
for dbg in 0 16 4 20; do
SGR_DEBUG=$dbg python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --gpu-reference 0 --train 0 > /tmp/b_$dbg.json 2>/dev/null
python - <<P
import json
d=json.loads(open('/tmp/b_$dbg.json').read().strip().splitlines()[-1])
print('debug $dbg: ms %.3f' % d['ms_per_step'], ' '.join('%s:%.3f' % (l.get('layer', i), l['ms']) for i, l in enumerate(d['layers'])))
P
done
