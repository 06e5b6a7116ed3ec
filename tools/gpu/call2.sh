#!/bin/bash
# round-2 GPU call 2: new parity tests, full-bench smoke with the new legs, ncu (source-level) of the scatter / halo kernels, HBM-kernel ncu
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/pytest_r2_c2.log
echo "== bench default"; (time python bench.py) > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err; tail -5 gpurun_out/bench_r2_c2.err; python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2_c2.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','sustained','strong_scaling','train_step','gpu_reference','cpu_baseline'):
        print(k, json.dumps(d.get(k))[:600])
    r=d['roofline']; print('roofline', {k:r[k] for k in ('achieved','frac','issued_frac','issued_frac_vs_burst','kernel_ms_per_step')})
    print('layers', [(l['layer'],l['ms']) for l in d['layers']])
except Exception as e: print('bench parse failed', e)
P
echo "== reference arm"; (time python bench.py --impl reference --steps 3 --warmup 1) 2>&1 | cut -c1-700 | tail -6
echo "== scatter L11 debug switches"
for d in 0 1 2; do SGR_DEBUG=$d python tools/gpu_layer_bench.py 32 L11 2>&1 | sed "s/^/DEBUG=$d /"; done
SGR_TMA_STORE=0 python tools/gpu_layer_bench.py 32 L11 2>&1 | sed "s/^/TMASTORE=0 /"
echo "== ncu conv kernels of one bench step (full set + source)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"modconv_halo_kernel|upconv_scatter" -s 33 -c 11 -o gpurun_out/prof_r2a python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ncu_r2a.log 2>&1; tail -2 gpurun_out/ncu_r2a.log
echo "== ncu hbm"; timeout 600 ncu --set full --clock-control none -k regex:"upfirdn2d_kernel|torgb_tail|bwd_act|up_bwd_prepare|param_sums|frames_to_uint8" -c 60 -o gpurun_out/prof_hbm python tools/gpu_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; tail -2 gpurun_out/ncu_hbm.log
ls -la gpurun_out/*.ncu-rep
