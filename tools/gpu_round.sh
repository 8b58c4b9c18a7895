#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/all_tests.log 2>&1; echo "all rc=$?"
tail -8 gpurun_out/all_tests.log
for fb in 1 0; do
  echo "== FUSED_BLUR=$fb"
  DSRG_B200_FUSED_BLUR=$fb timeout 300 python tools/bench_infer.py --cpu 0 > gpurun_out/infer_fb$fb.json 2> gpurun_out/infer.err; echo "infer rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/infer_fb$fb.json'))
print(d['value'], d['ms_per_image'], d['kernel_ms_per_image'], d['launches_per_image'], d['kernel_classes_ms_per_image'])
PY
  DSRG_B200_FUSED_BLUR=$fb timeout 600 python bench.py > gpurun_out/bench_fb$fb.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  python tools/bench_summary.py gpurun_out/bench_fb$fb.json 2>/dev/null | head -30
  DSRG_B200_FUSED_BLUR=$fb timeout 600 python bench.py --workload train41 > gpurun_out/bench41_fb$fb.json 2> gpurun_out/bench.err; echo "bench41 rc=$?"
  python tools/bench_summary.py gpurun_out/bench41_fb$fb.json 2>/dev/null | head -8
done
