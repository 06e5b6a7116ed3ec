#!/bin/bash
# round-2 GPU call 3: MMA-rate microbenchmark, ncu of the conv kernels of one bench step (kept < 64 MiB), HBM-kernel table
mkdir -p gpurun_out
echo "== mma microbench"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/microbench/mma_rate.cu && /tmp/mma_rate 2>&1 | tee gpurun_out/mma_rate.log
echo "== scatter box experiments (L11, L9)"
for box in "" "16,8,1" "14,9,1"; do for d in 0 2; do SGR_UP_BOX=$box SGR_DEBUG=$d python tools/gpu_layer_bench.py 32 "up 128" 2>&1 | sed "s/^/BOX=$box DEBUG=$d /"; done; done
for d in 0 2; do SGR_UP_HALO=0 SGR_DEBUG=$d python tools/gpu_layer_bench.py 32 "up 128" 2>&1 | sed "s/^/HALO=0 DEBUG=$d /"; done
echo "== ncu conv kernels of one bench step (full set + source)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"modconv_halo_kernel|upconv_scatter" -s 33 -c 11 -o /tmp/prof_r2a python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/ncu_r2a.log 2>&1; tail -2 gpurun_out/ncu_r2a.log
ncu -i /tmp/prof_r2a.ncu-rep --page raw --csv > gpurun_out/prof_r2a_raw.csv 2>/dev/null
cp /tmp/prof_r2a.ncu-rep gpurun_out/prof_r2a.ncu-rep
echo "== ncu hbm"; timeout 600 ncu --set full --clock-control none -k regex:"upfirdn2d_kernel|torgb_tail|bwd_act|up_bwd_prepare|param_sums|frames_to_uint8" -c 40 -o /tmp/prof_hbm python tools/gpu_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; tail -2 gpurun_out/ncu_hbm.log
python tools/ncu_summary.py hbm /tmp/prof_hbm.ncu-rep gpurun_out/r2_hbm_kernels_ncu.md
ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > gpurun_out/prof_hbm_raw.csv 2>/dev/null
du -sh gpurun_out
