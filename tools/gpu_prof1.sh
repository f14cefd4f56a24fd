#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_case.py 1920 32 f64 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_1920.csv python tools/prof_case.py 1920 32 f64 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"panel_factor|stage2_chase|rank_update|gemm_tn|gemm_nn" -s 40 -c 6 -o gpurun_out/prof_r1_a python tools/prof_case.py 1920 32 f64 1 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stage2_chase" -c 1 -o gpurun_out/prof_r1_s2 python tools/prof_case.py 1920 32 f64 1 s2 > gpurun_out/ncu3.log 2>&1
cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
