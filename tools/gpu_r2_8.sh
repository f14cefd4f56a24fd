#!/bin/bash
# round 2, call 8: full GPU suite, bench, panel epilogue breakdown, ncu launch list + full captures of the two hot kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu --durations=10 > gpurun_out/r2_t_all2.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_all2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
echo "bench rc=$?" >> gpurun_out/r2_bench3.err
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 600 python tools/panel_blk_timing.py 3840 32 f64 4096 64 f64 8192 64 f32 > gpurun_out/r2_blk_timing5.log 2>&1
# ncu (one tool per call): launch list of a short bench command, then the two hot kernels with the full set
python bench.py --steps 1 --warmup 3 --sizes 1920 --no-cpu-baseline > gpurun_out/r2_plain_bench1920.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench_n1920.csv \
    python bench.py --steps 1 --warmup 3 --sizes 1920 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
python tools/stage2_only.py 3840 32 f64 > gpurun_out/r2_plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stage2_fast_kernel -c 2 -f -o gpurun_out/prof_r2_s2fast \
    python tools/stage2_only.py 3840 32 f64 > gpurun_out/r2_ncu_s2.log 2>&1
python tools/stage1_only.py 3840 32 f64 > gpurun_out/r2_plain_s1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:panel_blk_kernel -s 30 -c 3 -f -o gpurun_out/prof_r2_panelblk \
    python tools/stage1_only.py 3840 32 f64 > gpurun_out/r2_ncu_s1.log 2>&1
tail -25 gpurun_out/r2_t_all2.log; tail -2 gpurun_out/r2_bench3.err; cat gpurun_out/r2_blk_timing5.log; tail -3 gpurun_out/r2_ncu_s2.log gpurun_out/r2_ncu_s1.log gpurun_out/r2_ncu_launch.log
