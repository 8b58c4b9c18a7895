#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_schedule.py "8,24,32" "8,12,20,24" "8,12,16,28" "6,10,16,32" "8,10,14,32" "4,8,12,16,24" "8,12,18,26" "8,16,40" "10,14,40" "8,12,44" "6,10,14,34" "8,12,20,12,12" 2>&1 | tee gpurun_out/sweep.log
DSRG_B200_DEBUG_TIMING=1 DSRG_B200_HOST_SCHEDULE="8,12,20,24" timeout 300 python tools/sweep_schedule.py "8,12,20,24" 2>&1 | grep "host pass" | tail -3
DSRG_B200_DEBUG_TIMING=1 timeout 300 python tools/sweep_schedule.py "8,24,32" 2>&1 | grep "host pass" | tail -3
