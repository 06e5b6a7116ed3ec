#!/bin/bash
# evidence call: full GPU tests, default bench, ncu launch list + full-set capture of the conv kernels (kept < 64 MiB)
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_r2_c19.log
echo "== bench default"; (time python bench.py) > gpurun_out/bench_r2_c19.json 2> gpurun_out/bench_r2_c19.err; tail -4 gpurun_out/bench_r2_c19.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2_c19.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.3f issued_frac %.3f (burst %.3f) kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['ms_per_step'], r['issued_frac'], r['issued_frac_vs_burst'], r['kernel_ms_per_step']), [l['ms'] for l in d['layers']])
print('hbm_pass', json.dumps(r['hbm_pass'])[:400]); print('sep', json.dumps(r.get('separate_fir_pass'))[:300])
print('gpu_reference', json.dumps(d['gpu_reference'])[:300]); print('cpu', d['cpu_baseline']); print('train', json.dumps(d['train_step'])[:300])
P
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > /dev/null 2>&1; wc -l gpurun_out/launches_r2.csv
echo "== ncu conv kernels (full set)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"modconv_halo_kernel|upconv_scatter|modconv_kernel|up_finish|splitk_finish" -s 51 -c 17 -o /tmp/prof_r2 python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ncu_r2.log 2>&1; tail -2 gpurun_out/ncu_r2.log
ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv 2>/dev/null; ls -la /tmp/prof_r2.ncu-rep; cp /tmp/prof_r2.ncu-rep gpurun_out/prof_r2.ncu-rep
echo "== ncu hbm kernels"; timeout 600 ncu --set full --clock-control none -k regex:"upfirdn2d|torgb_tail|bwd_act|up_bwd_prepare|param_sums|frames_to_uint8" -c 40 -o /tmp/prof_hbm python tools/gpu_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; tail -1 gpurun_out/ncu_hbm.log
python tools/ncu_summary.py hbm /tmp/prof_hbm.ncu-rep gpurun_out/r2_hbm_kernels_ncu.md
du -sh gpurun_out
