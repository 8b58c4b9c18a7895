#!/bin/bash
mkdir -p gpurun_out
L=$PWD/dsrg_b200/lib
timeout 900 python -m pytest tests/test_gpu_crf.py -q -m gpu -x > gpurun_out/all_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/all_tests.log
for v in prev base prev base; do
  if [ $v = base ]; then unset DSRG_B200_LIB; else export DSRG_B200_LIB=$L/libdsrg_b200_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench.err; echo "bench $v rc=$?"
  python tools/bench_summary.py gpurun_out/bench_$v.json 2>/dev/null | head -3
done
