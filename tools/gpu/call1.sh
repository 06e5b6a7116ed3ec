#!/bin/bash
# round-2 GPU call 1: reference CUDA path timing, scatter wrapped-halo A/B, tests, HBM-kernel ncu, sanitizer race/sync checks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv
echo "== layer bench (new: wrapped halo)"; python tools/gpu_layer_bench.py 32 2>&1 | tee gpurun_out/lb_r2_halo.log
echo "== layer bench (SGR_UP_HALO=0)"; SGR_UP_HALO=0 python tools/gpu_layer_bench.py 32 up 2>&1 | tee gpurun_out/lb_r2_nohalo.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_r2_c1.log
echo "== bench 200 steps"; python bench.py --steps 200 --warmup 5 --cpu-baseline 0 > gpurun_out/bench_r2_c1.json 2> gpurun_out/bench_r2_c1.err; cut -c1-300 gpurun_out/bench_r2_c1.json
echo "== bench 200 steps, SGR_UP_HALO=0"; SGR_UP_HALO=0 SGR_BENCH_CHILD=1 python bench.py --steps 200 --warmup 5 --cpu-baseline 0 > gpurun_out/bench_r2_c1_nohalo.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_c1_nohalo.json
echo "== reference on GPU"; (time python tools/gpu_reference_bench.py --steps 20) > gpurun_out/ref_gpu.json 2> gpurun_out/ref_gpu.err; tail -4 gpurun_out/ref_gpu.err; cat gpurun_out/ref_gpu.json
echo "== ncu hbm"; timeout 600 ncu --set full --clock-control none -k regex:"upfirdn2d_kernel|torgb_tail|bwd_act|up_bwd_prepare|param_sums|frames_to_uint8" -c 60 -o gpurun_out/prof_hbm python tools/gpu_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; tail -2 gpurun_out/ncu_hbm.log
echo "== racecheck"; timeout 400 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py 1 > gpurun_out/racecheck.log 2>&1; tail -5 gpurun_out/racecheck.log
echo "== synccheck"; timeout 300 compute-sanitizer --tool synccheck python tools/sanitizer_cases.py 1 > gpurun_out/synccheck.log 2>&1; tail -5 gpurun_out/synccheck.log
