#!/bin/bash
# quick device-resident numbers for kernel tuning (not reportable lines): tools/quick_bench.sh <tag> [workloads...]
tag=$1; shift
mkdir -p gpurun_out
for w in "${@:-dsrg321}"; do
  extra=""; steps=10
  case $w in noise) extra="--images noise"; wl=dsrg321; steps=3;; photo) extra="--images photo"; wl=dsrg321; steps=5;; *) wl=$w;; esac
  timeout 300 python bench.py --workload $wl $extra --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/q_${tag}_$w.json 2> gpurun_out/q_${tag}_$w.err || tail -3 gpurun_out/q_${tag}_$w.err
  echo "== $tag $w"; python tools/bench_summary.py gpurun_out/q_${tag}_$w.json 2>/dev/null | head -${QB_LINES:-7}
done
