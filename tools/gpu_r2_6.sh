#!/bin/bash
# round 2, call 6: panel kernel v3 + helper fix: timing breakdowns, parity, then the whole GPU suite and the bench
mkdir -p gpurun_out
export SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so
python tools/stage2_only.py 3840 32 f64 > gpurun_out/r2_s2f_phases2.log 2>&1
python tools/stage2_only.py 3840 32 f32 >> gpurun_out/r2_s2f_phases2.log 2>&1
timeout 600 python tools/panel_blk_timing.py 1920 32 f64 3840 32 f64 3840 32 f32 4096 64 f64 8192 64 f32 > gpurun_out/r2_blk_timing3.log 2>&1
unset SVDB200_LIB
timeout 900 python tools/panel_diag.py > gpurun_out/r2_panel_diag4.log 2>&1
timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r2_t_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
echo "bench rc=$?" >> gpurun_out/r2_bench2.err
cat gpurun_out/r2_s2f_phases2.log gpurun_out/r2_blk_timing3.log gpurun_out/r2_panel_diag4.log; tail -30 gpurun_out/r2_t_all.log; tail -3 gpurun_out/r2_bench2.err
