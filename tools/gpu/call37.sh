#!/bin/bash
# evidence call after the backward fusions + linear scatter tiles
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_r2_final.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; (time python bench.py --steps 20 --warmup 5) > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -4 gpurun_out/bench_r2_final.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f ms %.3f e2e %.0f u8 %.0f sustained %.3f issued_frac %.3f (burst %.3f) kernel_ms %.3f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_frames']['value'], d['sustained']['ms_per_step'], r['issued_frac'], r['issued_frac_vs_burst'], r['kernel_ms_per_step'], d['gpu_launches']), [round(l['ms'],3) for l in d['layers']])
print('hbm_pass', json.dumps(r['hbm_pass'])[:300]); print('sep', json.dumps(r.get('separate_fir_pass'))[:250])
print('gpu_reference', {k: d['gpu_reference'].get(k) for k in ('speedup_vs_reference_fp32','speedup_vs_reference_tf32_default','max_abs_ours_vs_reference_fp32')}); print('cpu', d['cpu_baseline']['value']); print('train', d['train_step']['ms_per_step'], d['train_step']['generator_only']['ms_per_step']); print('strong', d['strong_scaling']['value']); print('clocks', d['clocks'])
P
echo "== reference arm"; (time python bench.py --impl reference --steps 20 --warmup 5) 2>&1 | cut -c1-200 | tail -5
echo "== optimize_g step"; python tools/gpu_optimize_g_bench.py 1 2>&1 | tail -1; python tools/gpu_optimize_g_bench.py 16 2>&1 | tail -1
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > /dev/null 2>&1; wc -l gpurun_out/launches_r2.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/optg_b1_r2.csv python tools/gpu_optimize_g_bench.py 1 > /dev/null 2>&1; wc -l gpurun_out/optg_b1_r2.csv
echo "== ncu conv kernels (full set)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"modconv_halo_kernel|upconv_scatter|modconv_kernel|up_finish|splitk_finish" -s 48 -c 16 -o /tmp/prof_r2 python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ncu_r2.log 2>&1; tail -2 gpurun_out/ncu_r2.log
ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv 2>/dev/null; ls -la /tmp/prof_r2.ncu-rep; cp /tmp/prof_r2.ncu-rep gpurun_out/prof_r2.ncu-rep
echo "== ncu hbm kernels"; timeout 600 ncu --set full --clock-control none -k regex:"upfirdn2d|torgb_tail|bwd_act|up_bwd_prepare|frames_to_uint8" -c 40 -o /tmp/prof_hbm python tools/gpu_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; tail -1 gpurun_out/ncu_hbm.log
python tools/ncu_summary.py hbm /tmp/prof_hbm.ncu-rep gpurun_out/r2_hbm_kernels_ncu.md
du -sh gpurun_out
