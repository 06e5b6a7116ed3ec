#!/bin/bash
# end-of-round evidence: full GPU test suite, smoke, the default bench line, ncu launch list of a short bench run
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
(time python bench.py) > gpurun_out/bench.log 2>&1; tail -c 600 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --gpu-reference 0 --train 0 > gpurun_out/bench_ncu.log 2>&1; tail -c 200 gpurun_out/bench_ncu.log
