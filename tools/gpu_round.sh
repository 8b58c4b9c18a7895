#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/all_tests.log 2>&1; echo "all rc=$?"
tail -6 gpurun_out/all_tests.log
timeout 600 python -m pytest tests -q -m "not gpu" > gpurun_out/cpu_tests.log 2>&1; echo "cpu rc=$?"
tail -3 gpurun_out/cpu_tests.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-700 gpurun_out/bench_ref.json
