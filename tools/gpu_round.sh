#!/bin/bash
mkdir -p gpurun_out
DSRG_B200_TWIN=1 timeout 300 python -m pytest tests/test_gpu_srg.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -2
DSRG_B200_TWIN=1 DSRG_B200_DEBUG_TIMING=1 timeout 200 python tools/sweep_schedule.py "8,24,32" "8,16,40" "12,20,32" "8,24,32" 2>&1 | grep -v "^\[dsrg host pass\] total" | tail -12
