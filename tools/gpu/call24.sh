#!/bin/bash
mkdir -p gpurun_out
echo "== bench N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; tail -2 gpurun_out/bench_r2_n8.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2_n8.json').read().strip().splitlines()[-1])
print('N=8 value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['value']))
print('strong', d['strong_scaling']['value'], d['strong_scaling']['ms']); print('train', json.dumps(d['train_step'])[:800])
P
echo "== config 1024 N=8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --config 1024 > gpurun_out/bench_r2_1024_n8.json 2> gpurun_out/bench_r2_1024_n8.err; tail -2 gpurun_out/bench_r2_1024_n8.err | cut -c1-300
echo "== config 1024 N=1"
timeout 900 python bench.py --steps 20 --config 1024 > gpurun_out/bench_r2_1024_n1.json 2> gpurun_out/bench_r2_1024_n1.err
python - <<'P'
import json
for n in (1, 8):
    try:
        d=json.loads(open('gpurun_out/bench_r2_1024_n%d.json' % n).read().strip().splitlines()[-1])
        print('1024 N=%d best %.0f frames/s' % (n, d['value']), [(r['batch_per_gpu'], round(r['bf16']['frames_s']), round(r['bf16x3']['frames_s'])) for r in d['sweep']], d['parity'])
    except Exception as e: print('1024 N=%d failed' % n, e)
P
