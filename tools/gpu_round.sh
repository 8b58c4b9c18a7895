#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/all_tests.log 2>&1; echo "all rc=$?"
tail -25 gpurun_out/all_tests.log
