#!/bin/bash
mkdir -p gpurun_out
echo "== latency"; python tools/gpu_latency_check.py 2>&1 | tee gpurun_out/latency_r2.log
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_cases.py 0 1 > gpurun_out/sanitizer_memcheck_r2.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_r2.log
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py 1 > gpurun_out/sanitizer_racecheck_r2.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_r2.log
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck python tools/sanitizer_cases.py 1 > gpurun_out/sanitizer_synccheck_r2.log 2>&1; tail -3 gpurun_out/sanitizer_synccheck_r2.log
