#!/bin/bash
# one gpurun call: new post-processing tests, then the whole GPU suite, inference latency, headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post.py -x -q -m gpu > gpurun_out/post_tests.log 2>&1; echo "post rc=$?" 
tail -15 gpurun_out/post_tests.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/all_tests.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/all_tests.log
timeout 300 python tools/bench_infer.py > gpurun_out/infer.json 2> gpurun_out/infer.err; echo "infer rc=$?"
cat gpurun_out/infer.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-600
