#!/bin/bash
# One gpurun call that validates a build and captures the evidence kept under profiles/:
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r2z'
# then, back in the dev container:
#   python tools/make_profiles.py r2z gpurun_out/bench_r2z.json gpurun_out/launches_r2z.csv \
#          gpurun_out/prof_r2z_tile.ncu-rep gpurun_out/bench_ref_r2z.json
T=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/all_tests_$T.log 2>&1; echo "tests rc=$?"
grep "graph replay\|passed\|failed" gpurun_out/all_tests_$T.log | tail -4
# the replay test is statistical (float atomics): run it a few more times and keep its diagnostic line
[ "${REPLAY_LOOP:-1}" = 1 ] && for i in 1 2 3 4; do timeout 100 python -m pytest tests/test_gpu_srg.py -q -m gpu -s -k graph_replay 2>&1 | grep "graph replay\|failed\|Error" ; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/bench_$T.json 2>/dev/null | head -16
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
# the other BASELINE configurations, each with its cpu_baseline leg
for w in crf321 srg321 full513 train41 sweep4096; do
  st=10; [ $w = sweep4096 ] && st=2
  timeout 600 python bench.py --workload $w --steps $st --warmup 3 > gpurun_out/bench_${T}_$w.json 2> gpurun_out/bench_${T}_$w.err; echo "$w rc=$?"
done
# launch list of exactly one step (plain launches: 131 per step -- 119 + 11 k_mf_tile_hy + k_tile_demote; the first 131 are a warm-up step)
DSRG_B200_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 131 -c 131 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
if [ "${NCU_FULL:-1}" = 1 ]; then
# full metric set: the tile kernel (skip the warm-up step's 11 launches and the FIRST-mode launch), then the rest
DSRG_B200_GRAPHS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^k_mf_tile$" -s 12 -c 2 -o gpurun_out/prof_${T}_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_t.log 2>&1
# the hybrid tiles' kernel on textured images (second step, a MID launch)
DSRG_B200_GRAPHS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^k_mf_tile_hy$" -s 12 -c 1 -o gpurun_out/prof_${T}_tile_hy \
    python bench.py --images photo --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_h.log 2>&1
fi
timeout 200 python tools/bench_infer.py > gpurun_out/infer_$T.json 2>/dev/null
ls -la gpurun_out/*$T*.ncu-rep gpurun_out/launches_$T.csv 2>/dev/null
