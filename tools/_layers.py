"""stdin: bench.py JSON line -> step ms + per-layer ms (helper for A/B runs)."""
import json
import sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.3f ms/step  %.0f frames/s  layers %s  fir %.3f' % (d['ms_per_step'], d['value'], [l['ms'] for l in d['layers']],
                                                          d['roofline']['hbm_pass']['ms_per_step']))
