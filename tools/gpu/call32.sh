#!/bin/bash
python tools/gpu_optimize_g_cpuprofile.py 1 2>&1 | grep -v "^$" | cut -c1-200 | head -90
