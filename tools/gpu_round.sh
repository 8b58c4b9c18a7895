#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_mf_tile" -s 12 -c 2 -o gpurun_out/prof_r1z_tile python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_t.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_lattice_insert<5>|k_tile_build<6|k_srg_label|k_srg_merge|k_srg_emit|k_norm_splat" -s 6 -c 6 -o gpurun_out/prof_r1z_misc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_m.log 2>&1
ls -la gpurun_out/*.ncu-rep
