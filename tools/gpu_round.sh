#!/bin/bash
mkdir -p gpurun_out
L=$PWD/dsrg_b200/lib
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/all_tests.log 2>&1; echo "all rc=$?"
tail -3 gpurun_out/all_tests.log
for v in pad0 base pad0 base; do
  if [ $v = base ]; then unset DSRG_B200_LIB; else export DSRG_B200_LIB=$L/libdsrg_b200_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench.err; echo "bench $v rc=$?"
  python tools/bench_summary.py gpurun_out/bench_$v.json 2>/dev/null | head -3
done
unset DSRG_B200_LIB
timeout 300 python bench.py --workload train41 --no-cpu-baseline > gpurun_out/bench41.json 2> gpurun_out/bench.err; python tools/bench_summary.py gpurun_out/bench41.json | head -2
timeout 300 python bench.py --images noise --no-cpu-baseline > gpurun_out/bench_noise.json 2> gpurun_out/bench.err; python tools/bench_summary.py gpurun_out/bench_noise.json | head -2
