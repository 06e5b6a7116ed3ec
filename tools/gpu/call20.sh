#!/bin/bash
mkdir -p gpurun_out
echo "== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or parity or modconv or nonseparable" 2>&1 | tail -2
echo "== layer bench up"; python tools/gpu_layer_bench.py 32 up 2>&1
echo "== bench N=2 (torchrun)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -3 gpurun_out/bench_r2_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['value']))
print('strong', json.dumps(d['strong_scaling'])[:300]); print('train', json.dumps(d['train_step'])[:700])
P
echo "== reference arm N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
