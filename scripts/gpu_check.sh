#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_reference_app.py tests/test_gpu_sim.py -q -s > gpurun_out/refapp.log 2>&1
grep -n "reference app\|passed\|failed" gpurun_out/refapp.log
