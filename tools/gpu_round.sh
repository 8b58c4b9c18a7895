#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/all_tests.log 2>&1; echo "all rc=$?"
tail -6 gpurun_out/all_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
LB=$PWD/dsrg_b200/lib/libdsrg_b200_lb.so
DSRG_B200_LIB=$LB timeout 600 python -m pytest tests/test_gpu_crf.py tests/test_gpu_dropin.py -q -m gpu -x > gpurun_out/lb_tests.log 2>&1; echo "lb tests rc=$?"
tail -4 gpurun_out/lb_tests.log
for v in base lb base lb; do
  if [ $v = lb ]; then export DSRG_B200_LIB=$LB; else unset DSRG_B200_LIB; fi
  timeout 600 python bench.py > gpurun_out/bench_$v.json 2> gpurun_out/bench.err; echo "bench $v rc=$?"
  python tools/bench_summary.py gpurun_out/bench_$v.json 2>/dev/null | head -5
done
unset DSRG_B200_LIB
