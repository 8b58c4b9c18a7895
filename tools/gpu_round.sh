#!/bin/bash
# One gpurun call that validates a build and captures the evidence kept under profiles/:
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'
# then, back in the dev container:
#   python tools/make_profiles.py r1z gpurun_out/bench_r1z.json gpurun_out/launches_r1z.csv \
#          gpurun_out/prof_r1z_tile.ncu-rep gpurun_out/bench_ref_r1z.json
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/all_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1z.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/bench_r1z.json 2>/dev/null | head -16
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1z.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
# launch list of exactly one step (119 launches per step; the first 119 are the warm-up step)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 119 -c 119 --csv --log-file gpurun_out/launches_r1z.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
# full metric set: the tile kernel (skip the warm-up step's 11 launches and the FIRST-mode launch), then the rest
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_mf_tile" -s 12 -c 2 -o gpurun_out/prof_r1z_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_t.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_mf_blur|k_lattice_insert<5>|k_tile_build<6|k_srg_label|k_srg_merge|k_srg_emit|k_norm_splat" \
    -s 100 -c 8 -o gpurun_out/prof_r1z_misc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_m.log 2>&1
timeout 200 python tools/bench_infer.py > gpurun_out/infer.json 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1z.csv
