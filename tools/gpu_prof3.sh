#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_case.py 3840 32 f64 1 s1 > gpurun_out/prof3_plain.log 2>&1 &&
SVDB200_LOOKAHEAD=0 ncu --set full --clock-control none --import-source on -k regex:"panel_reg" -s 20 -c 2 -o gpurun_out/prof_r1_panelreg python tools/prof_case.py 3840 32 f64 1 s1 > gpurun_out/ncu_p3.log 2>&1
cat gpurun_out/prof3_plain.log; tail -n 3 gpurun_out/ncu_p3.log
