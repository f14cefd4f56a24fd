#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_case.py 3840 32 f64 1 s1 > gpurun_out/prof2_plain.log 2>&1 &&
SVDB200_LOOKAHEAD=0 ncu --set full --clock-control none --import-source on -k regex:"panel_factor" -s 20 -c 2 -o gpurun_out/prof_r1_panel python tools/prof_case.py 3840 32 f64 1 s1 > gpurun_out/ncu_p.log 2>&1
SVDB200_LOOKAHEAD=0 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_fast|gemm_nn_fast|rank_update_fast" -s 6 -c 3 -o gpurun_out/prof_r1_gemm16k python tools/prof_case.py 16384 64 f64 1 s1 > gpurun_out/ncu_g.log 2>&1
cat gpurun_out/prof2_plain.log; tail -n 3 gpurun_out/ncu_p.log gpurun_out/ncu_g.log
