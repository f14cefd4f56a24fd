#!/bin/bash
# round 2, call 23: full GPU suite, bench (both arms), ncu launch list of the bench command, full captures of the Cholesky-QR panel
# kernels (stage 1 at n = 16384, band 64, float) and of the stage-2 kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2_t_all3.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_all3.log
tail -14 gpurun_out/r2_t_all3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
echo "bench rc=$?" >> gpurun_out/r2_bench4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench4_ref.json 2> gpurun_out/r2_bench4_ref.err
echo "ref rc=$?" >> gpurun_out/r2_bench4_ref.err
tail -2 gpurun_out/r2_bench4.err gpurun_out/r2_bench4_ref.err
python bench.py --steps 1 --warmup 3 --sizes 1920 --no-cpu-baseline > gpurun_out/r2_plain_bench1920b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench_n1920b.csv \
    python bench.py --steps 1 --warmup 3 --sizes 1920 --no-cpu-baseline > gpurun_out/r2_ncu_launchb.log 2>&1
REPS=2 python tools/stage1_only.py 16384 64 f32 > gpurun_out/r2_plain_s1b.log 2>&1 && \
REPS=1 ncu --set full --clock-control none --import-source on -k regex:"chol_(gram|algebra|apply)_kernel" -s 60 -c 6 -f -o gpurun_out/prof_r2_chol \
    python tools/stage1_only.py 16384 64 f32 > gpurun_out/r2_ncu_s1b.log 2>&1
tail -2 gpurun_out/r2_plain_s1b.log gpurun_out/r2_ncu_s1b.log gpurun_out/r2_ncu_launchb.log
