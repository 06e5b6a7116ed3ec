#!/bin/bash
python tools/gpu_optimize_g_cpuprofile.py 1 2>&1 | grep -E "^sync|^plain"
