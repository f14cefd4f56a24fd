#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batched" -p no:cacheprovider 2>&1 | tail -5
timeout 300 python tools/prof_case.py 1920 32 f64 1 s2 > gpurun_out/prof4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stage2_chase" -c 1 -o gpurun_out/prof_r1_s2b python tools/prof_case.py 1920 32 f64 1 s2 > gpurun_out/ncu_s2b.log 2>&1
cat gpurun_out/prof4_plain.log; tail -n 2 gpurun_out/ncu_s2b.log
